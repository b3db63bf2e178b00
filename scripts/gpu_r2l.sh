# round 2, GPU call L (2 GPUs): final build sanity — sharded / peer all-reduce tests, full single-GPU suite, bench at 2 GPUs
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2l.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 20 --warmup 3 --c5-iters 1 > gpurun_out/bench_r2l_g2.json 2> gpurun_out/bench_r2l_g2.err; echo "bench rc=$?" >> gpurun_out/bench_r2l_g2.err
echo done
