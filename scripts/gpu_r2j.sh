# round 2, GPU call J: probe T9 (A operand from TMEM); E-step with the event operand in TMEM (default build) against the
# shared-memory version (variants/libgmm_b200_v2.so): parity subset under a timeout, A/B timing, then the full suite
set -x
mkdir -p gpurun_out
timeout 120 cuda-gmm-mpi_b200/csrc/probe/tc_probe > gpurun_out/probe_r2j.log 2>&1
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "estep_parity or tensor_path_large or tight or rerun" > gpurun_out/pytest_r2j_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2j_quick.log
export GMM_EXP_N=4000000
timeout 300 python scripts/exp_ab.py default cuda-gmm-mpi_b200/variants/libgmm_b200_v2.so > gpurun_out/ab_r2j.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2j.log
echo done
