set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu5.log
timeout 240 python bench.py --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/bench_c3_d.json 2> gpurun_out/bench_c3_d.err
echo done
