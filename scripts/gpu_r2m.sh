# round 2, GPU call M: E-step without the duplicated zh chunks in the A image (default build) against the previous build
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "estep_parity or tensor_path_large or tight or rerun or cli" > gpurun_out/pytest_r2m_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2m_quick.log
export GMM_EXP_N=4000000
timeout 300 python scripts/exp_ab.py default cuda-gmm-mpi_b200/variants/libgmm_b200_base.so default cuda-gmm-mpi_b200/variants/libgmm_b200_base.so > gpurun_out/ab_r2m.log 2>&1
echo done
