set -x
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q -x -k "estep" > gpurun_out/pytest_gpu6a.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_gpu6a.log
if [ $rc -ne 0 ]; then echo "estep tests failed; skipping bench"; exit 0; fi
timeout 240 python bench.py --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/bench_c3_e.json 2> gpurun_out/bench_c3_e.err
timeout 120 python bench.py --workload c2 --steps 20 --warmup 3 --cpu-sample 0 > gpurun_out/bench_c2_e.json 2> gpurun_out/bench_c2_e.err
echo done
