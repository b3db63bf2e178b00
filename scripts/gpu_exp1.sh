set -x
mkdir -p gpurun_out
: > gpurun_out/exp_mstep_acc.log
for lib in "" cuda-gmm-mpi_b200/variants/libgmm_b200_c1_s8.so cuda-gmm-mpi_b200/variants/libgmm_b200_c2_s8.so; do
  GMM_B200_LIB=$lib timeout 200 python scripts/exp_mstep_acc.py >> gpurun_out/exp_mstep_acc.log 2>&1
  GMM_B200_LIB=$lib timeout 200 python bench.py --steps 5 --warmup 3 --no-e2e --cpu-sample 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['phases_ms_per_step'])" >> gpurun_out/exp_mstep_acc.log 2>&1
done
echo done
