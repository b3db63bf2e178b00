# round 2, GPU call O (2 GPUs): device-side finalisation across ranks — sharded C-ABI parity, finalisation parity/replay, bench at N = 2 and 1
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scale.py -m gpu -q -x -k "device_finalisation or sharded or peer_memory or em_iterations" > gpurun_out/pytest_r2o.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2o.log
for M in device host; do
  GMM_FINALIZE=$M timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 2 --steps 20 --warmup 3 --repeats 3 --no-ref-gpu --cpu-sample 0 --c5-iters 0 > gpurun_out/bench_r2o_g2_$M.json 2> gpurun_out/bench_r2o_g2_$M.err
done
GMM_FINALIZE=device timeout 300 python bench.py --steps 20 --warmup 3 --repeats 3 --no-ref-gpu --cpu-sample 0 --c5-iters 0 > gpurun_out/bench_r2o_g1_device.json 2> gpurun_out/bench_r2o_g1_device.err
echo done
