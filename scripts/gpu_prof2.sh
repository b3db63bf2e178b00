set -x
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"estep_tc_kernel|mstep_tc_kernel" -s 4 -c 2 -o gpurun_out/prof_c3_r1b python bench.py --steps 3 --warmup 3 --no-e2e --cpu-sample 0 > gpurun_out/ncu_full2.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/bench_c3_f.json 2> gpurun_out/bench_c3_f.err
timeout 600 python -m pytest tests -m gpu -q -k "mstep" > gpurun_out/pytest_gpu7.log 2>&1
echo done
