# round 2, GPU call I (2 GPUs): peer-memory all-reduce kernel vs NCCL, sharded parity, bench at 2 GPUs with both
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "sharded or two_gpus or peer_memory" > gpurun_out/pytest_r2i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2i.log
for AR in 1 0; do
  GMM_BENCH_ALLREDUCE=$AR timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2963$AR bench.py --gpus 2 --steps 20 --warmup 3 --c5-iters 0 > gpurun_out/bench_r2i_ar$AR.json 2> gpurun_out/bench_r2i_ar$AR.err; echo "bench rc=$?" >> gpurun_out/bench_r2i_ar$AR.err
done
echo done
