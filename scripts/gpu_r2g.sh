# round 2, GPU call G: E-step stage-isolation timing variants; racecheck at the production shape
set -x
mkdir -p gpurun_out
export GMM_EXP_N=4000000
V=cuda-gmm-mpi_b200/variants
timeout 400 python scripts/exp_ab.py default $V/libgmm_b200_ex1.so $V/libgmm_b200_ex2.so $V/libgmm_b200_ex3.so > gpurun_out/ab_r2g.log 2>&1
SAN_N=1500 SAN_D=24 SAN_K=64 timeout 600 compute-sanitizer --tool racecheck python scripts/san_run.py > gpurun_out/san_race24_r2g.log 2>&1
echo done
