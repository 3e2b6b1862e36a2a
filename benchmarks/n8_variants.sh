# data-parallel variants of bench.py on N GPUs of one node: SMs reserved for the collective while it overlaps backward
N=${N:-8}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/n${N}_$2.err | tee gpurun_out/bench_n${N}_$2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2', d['value'], d['ms_per_step'], d.get('comm'))"; }
LG_DP_RESERVE_SMS=0 run 29521 reserve0
LG_DP_RESERVE_SMS=8 run 29522 reserve8
LG_DP_RESERVE_SMS=16 run 29523 reserve16
LG_DP_RESERVE_SMS=32 run 29524 reserve32
