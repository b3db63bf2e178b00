set -x
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"estep_tc_kernel" -s 2 -c 1 -o gpurun_out/prof_c3_r1c python bench.py --steps 2 --warmup 3 --no-e2e --cpu-sample 0 > gpurun_out/ncu_full3.log 2>&1
echo done
