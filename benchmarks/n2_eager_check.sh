# does the gradient exchange overlap backward once some SMs are really left free for NCCL?
N=${N:-2}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 6 --warmup 3 $3 2>gpurun_out/n${N}_$2.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2', d['value'], d['ms_per_step'], d.get('comm'))"; }
LG_NO_SIDE_STREAM=1 LG_DP_RESERVE_SMS=16 run 29531 noside_reserve16
LG_NO_SIDE_STREAM=1 LG_DP_RESERVE_SMS=32 run 29532 noside_reserve32
LG_NO_SIDE_STREAM=1 run 29533 noside_reserve0
