# usage (from the repo root):  gpurun --gpus N -- 'N=<N> bash scripts/gpu_multiN.sh'
# sharded C-ABI parity tests (self-skipping above the visible GPU count) + bench.py under torchrun at every power of two up to N
set -x
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_g${N}.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "sharded or two_gpus" > gpurun_out/pytest_multi_g${N}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi_g${N}.log
for G in ${GLIST:-$N}; do
  NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2961$G bench.py --gpus $G --steps 20 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_r2_g${G}.json 2> gpurun_out/bench_r2_g${G}.err; echo "bench rc=$?" >> gpurun_out/bench_r2_g${G}.err
done
echo done
