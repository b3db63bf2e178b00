# round 2, GPU call H: packed-f32x2 log-sum-exp stage, one-factorisation host finalisation, worker pool: full suite + timing
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_r2h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2h.log
timeout 600 python bench.py --steps 20 --warmup 3 --repeats 5 --no-ref-gpu --cpu-sample 0 --c5-iters 2 > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; echo "bench rc=$?" >> gpurun_out/bench_r2h.err
echo done
