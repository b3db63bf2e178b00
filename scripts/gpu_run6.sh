set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x -k "estep" > gpurun_out/pytest_gpu6a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu6a.log
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/bench_c3_e.json 2> gpurun_out/bench_c3_e.err
timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 --cpu-sample 0 > gpurun_out/bench_c2_e.json 2> gpurun_out/bench_c2_e.err
echo done
