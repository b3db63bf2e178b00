set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "estep or em_config or fit or smoke" > gpurun_out/pytest_exp3.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/pytest_exp3.log
if [ $rc -eq 0 ]; then timeout 300 python bench.py > gpurun_out/bench_exp3.json 2> gpurun_out/bench_exp3.err; echo "rc=$?" >> gpurun_out/bench_exp3.err; fi
echo done
