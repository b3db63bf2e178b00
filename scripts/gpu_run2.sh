set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/bench_c3_b.json 2> gpurun_out/bench_c3_b.err
echo done
