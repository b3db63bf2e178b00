# round 2, GPU call D: full suite with the restored run-level bars (no -x: every deviation is reported), bench.py, ncu of both kernels
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_r2d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2d.log
timeout 900 python bench.py --steps 10 --warmup 3 --repeats 3 --c5-iters 1 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo "bench rc=$?" >> gpurun_out/bench_r2d.err
export GMM_EXP_N=4000000
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"estep_tc_kernel|mstep_tc_kernel" -s 4 -c 2 -f -o gpurun_out/prof_r2d python scripts/prof_run.py > gpurun_out/ncu_r2d.log 2>&1
echo done
