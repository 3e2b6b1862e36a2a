# does the gradient exchange overlap backward when GEMM CTAs leave shared memory for NCCL's CTAs?
# (variant library: python -c "from lightgrad_b200 import build as b; b.build(extra_flags=['-DLG_GEMM_SMEM_CUT_KB=48'],
#  lib='lightgrad_b200/lib/variant_smemcut48.so', obj_dir='lightgrad_b200/csrc/build_cut48')"; result: no, see lg_gemm_tc.cu)
N=${N:-2}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 6 --warmup 3 $3 2>gpurun_out/n${N}_$2.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2', d['value'], d['ms_per_step'], d.get('comm'))" | tee -a gpurun_out/n${N}_variants.txt; }
LG_LIB=$PWD/lightgrad_b200/lib/variant_smemcut48.so run 29531 smemcut48
