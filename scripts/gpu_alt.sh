# default path after the ALT refactor + first correctness / timing reading of GMM_ESTEP_ALT=1
set -x
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "estep_tensor_path_large or em_config1" > gpurun_out/pytest_alt_default.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_alt_default.log
GMM_ESTEP_ALT=1 timeout 45 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "estep_tensor_path_large" > gpurun_out/pytest_alt_on.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_alt_on.log
GMM_ESTEP_ALT=1 timeout 60 python scripts/exp_ab.py default > gpurun_out/exp_ab_alt.log 2>&1
echo done
