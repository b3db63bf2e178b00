# round 2, GPU call A: full GPU suite on the RN-split build, probe T7/T8, error budget of the tensor path, E-step variants
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_r2a.txt; nproc >> gpurun_out/gpus_r2a.txt
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_r2a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2a.log
timeout 120 cuda-gmm-mpi_b200/csrc/probe/tc_probe > gpurun_out/probe_r2a.log 2>&1
V=cuda-gmm-mpi_b200/variants
timeout 1500 python scripts/exp_acc.py all default $V/libgmm_b200_tr.so $V/libgmm_b200_p4.so $V/libgmm_b200_c2.so $V/libgmm_b200_c1.so $V/libgmm_b200_trc1.so $V/libgmm_b200_nst4.so > gpurun_out/exp_acc_r2a.log 2>&1
export GMM_EXP_N=4000000
timeout 200 python scripts/exp_ab.py $V/libgmm_b200_eprof.so > gpurun_out/eprof_r2a.log 2>&1
GMM_ESTEP_WG4=1 timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k estep_tensor > gpurun_out/pytest_wg4_r2a.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_wg4_r2a.log
GMM_ESTEP_WG4=1 timeout 200 python scripts/exp_ab.py default > gpurun_out/ab_wg4_r2a.log 2>&1
GMM_ESTEP_WG4=1 timeout 200 python scripts/exp_ab.py $V/libgmm_b200_eprof.so > gpurun_out/eprof_wg4_r2a.log 2>&1
echo done
