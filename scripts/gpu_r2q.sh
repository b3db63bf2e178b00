# round 2, GPU call Q: finalize_params_kernel with one barrier per column / row (short dependent chains) — parity, replay, timing
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_scale.py -m gpu -q -x -k "device_finalisation or em_iterations" > gpurun_out/pytest_r2q.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2q.log
timeout 300 python bench.py --steps 20 --warmup 3 --repeats 3 --no-ref-gpu --cpu-sample 0 --c5-iters 0 > gpurun_out/bench_r2q.json 2> gpurun_out/bench_r2q.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:finalize_params -c 6 --csv --log-file gpurun_out/launches_fin_r2q.csv python scripts/prof_run.py > gpurun_out/ncu_r2q.log 2>&1
echo done
