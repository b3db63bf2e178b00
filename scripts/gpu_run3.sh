set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
echo done
