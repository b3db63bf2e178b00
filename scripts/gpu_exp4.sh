set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mstep or em_config or fit or upload or convergence or full_size" > gpurun_out/pytest_exp4.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/pytest_exp4.log
timeout 300 python bench.py > gpurun_out/bench_exp4.json 2> gpurun_out/bench_exp4.err
echo done
