# data-parallel variants of bench.py on N GPUs of one node (N=8 by default): optimizer pipelined behind the
# bucketed all-reduce (LG_DP_PIPELINED_STEP=1) vs plain optimizer.step() after backward (default)
N=${N:-8}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/n${N}_$2.err | tee gpurun_out/bench_n${N}_$2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2', d['value'], d['ms_per_step'], d['loss'], d['e2e']['last_loss'], d.get('comm'))"; }
LG_DP_PIPELINED_STEP=1 run 29521 pipelined
run 29522 plainstep
