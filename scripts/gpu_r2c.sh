# round 2, GPU call C: M-step with gs remainder product + rare remainder drains; tests, error budget, first bench.py run
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_r2c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2c.log
timeout 600 python scripts/exp_acc.py all default > gpurun_out/exp_acc_r2c.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --repeats 3 --c5-iters 1 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench rc=$?" >> gpurun_out/bench_r2c.err
echo done
