# round 2, GPU call N: device-side finalisation (finalize_params_kernel) — parity against the host finalisation and the replay,
# then the whole GPU suite, then bench.py with GMM_FINALIZE=device (default) and =host
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_scale.py -m gpu -q -x -k "device_finalisation" > gpurun_out/pytest_r2n_fin.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2n_fin.log
for M in device host; do
  GMM_FINALIZE=$M timeout 300 python bench.py --steps 20 --warmup 3 --repeats 3 --no-ref-gpu --cpu-sample 0 --c5-iters 0 > gpurun_out/bench_r2n_$M.json 2> gpurun_out/bench_r2n_$M.err
done
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_r2n_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2n_all.log
echo done
