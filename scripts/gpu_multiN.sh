set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_c3_g2_b.json 2> gpurun_out/bench_c3_g2_b.err
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "two_gpus" > gpurun_out/pytest_multi2_b.log 2>&1
echo done
