# round 2, GPU call B: new M-step (exact fixed-point leading products, N=128 stacked B, single-buffered staggered flush),
# scaled E-step operand, new tests; error budget + timing; E-step epilogue phase profile
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_r2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2b.log
timeout 600 python scripts/exp_acc.py all default > gpurun_out/exp_acc_r2b.log 2>&1
export GMM_EXP_N=4000000
timeout 200 python scripts/exp_ab.py cuda-gmm-mpi_b200/variants/libgmm_b200_eprof.so > gpurun_out/eprof_r2b.log 2>&1
echo done
