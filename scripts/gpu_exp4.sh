set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gamma_modes or em_config2 or full_size" > gpurun_out/pytest_exp4.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/pytest_exp4.log
timeout 200 python bench.py --no-e2e --cpu-sample 0 > gpurun_out/bench_exp4.json 2> gpurun_out/bench_exp4.err
echo done
