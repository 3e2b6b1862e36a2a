# data-parallel variants of bench.py on N GPUs of one node (N=8 by default); results of round 1 in DESIGN.md section 6
N=${N:-8}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/n${N}_$2.err | tee gpurun_out/bench_n${N}_$2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2', d['value'], d['ms_per_step'], d['loss'], d.get('comm'))" | tee -a gpurun_out/n${N}_variants.txt; }
run 29521 default
LG_DP_NO_OVERLAP=1 run 29522 blocking_allreduce
LG_DP_PIPELINED_STEP=1 run 29523 pipelined_adam
# SMs really left free for a collective capped to as many CTAs (no side stream: it would backfill them)
LG_NO_SIDE_STREAM=1 LG_DP_RESERVE_SMS=16 NCCL_MAX_CTAS=16 run 29524 noside_reserve16_ctas16
