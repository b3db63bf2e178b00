set -x
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c3_r1e.csv python bench.py --steps 3 --warmup 3 --no-e2e --cpu-sample 0 > gpurun_out/ncu_launch5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"estep_tc_kernel|mstep_tc_kernel" -s 4 -c 2 -o gpurun_out/prof_c3_r1e python bench.py --steps 3 --warmup 3 --no-e2e --cpu-sample 0 > gpurun_out/ncu_full5.log 2>&1
echo done
