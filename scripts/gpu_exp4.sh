set -x
mkdir -p gpurun_out
timeout 120 python bench.py --workload c2 --steps 50 --warmup 3 --no-e2e --cpu-sample 0 > gpurun_out/bench_c2_f.json 2> gpurun_out/bench_c2_f.err
echo done
