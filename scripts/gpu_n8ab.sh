# round 2, 8-GPU call 2: the per-iteration all-reduce as the library's peer-memory kernel (default) against ncclAllReduce, bench.py at N = 8
set -x
mkdir -p gpurun_out
for A in 1 0; do
  GMM_BENCH_ALLREDUCE=$A NCCL_DEBUG=WARN timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2963$A bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e --no-ref-gpu --cpu-sample 0 --c5-iters 0 > gpurun_out/bench_n8_allreduce$A.json 2> gpurun_out/bench_n8_allreduce$A.err
done
echo done
