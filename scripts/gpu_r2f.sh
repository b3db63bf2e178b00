# round 2, GPU call F: ncu of the new kernels, K>64 single-write E-step, rerun-stability stress, sanitizers
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_path_large or config5 or rerun or estep_parity" > gpurun_out/pytest_r2f_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2f_quick.log
export GMM_EXP_N=4000000
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"estep_tc_kernel|mstep_tc_kernel" -s 4 -c 2 -f -o gpurun_out/prof_r2f python scripts/prof_run.py > gpurun_out/ncu_r2f.log 2>&1
SAN_N=6000 timeout 300 compute-sanitizer --tool memcheck python scripts/san_run.py > gpurun_out/san_mem_r2f.log 2>&1
SAN_N=3000 timeout 300 compute-sanitizer --tool synccheck python scripts/san_run.py > gpurun_out/san_sync_r2f.log 2>&1
SAN_N=2000 SAN_D=8 SAN_K=6 timeout 400 compute-sanitizer --tool racecheck python scripts/san_run.py > gpurun_out/san_race_r2f.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --repeats 5 --no-ref-gpu --cpu-sample 0 > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; echo "bench rc=$?" >> gpurun_out/bench_r2f.err
echo done
