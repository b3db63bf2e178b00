set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1; nproc >> gpurun_out/gpus.txt; free -g >> gpurun_out/gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 bash tests/golden/make_ref_golden.sh gpurun_out/golden > gpurun_out/golden.log 2>&1
timeout 600 python bench.py --workload c2 --steps 20 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
# reference program on B200 (unmodified kernels), config 2, 3 iterations
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import __graft_entry__ as e
pkg = e.load_package()
ev = pkg.synth.make_blobs(1_000_000, 16, 32)
pkg.synth.write_bin("/tmp/c2.bin", ev)
PY
OMP_NUM_THREADS=1 CUDA_VISIBLE_DEVICES=0 GMM_REF_ITERS=3 timeout 600 ./oracle/_ref/gaussianMPI_ref 32 /tmp/c2.bin /tmp/refc2 32 > gpurun_out/ref_on_b200_c2.log 2>&1
rm -f /tmp/refc2.results
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_simt_c2.csv python bench.py --workload c2 --steps 3 --warmup 3 --no-e2e --cpu-sample 0 > gpurun_out/ncu_bench.log 2>&1
echo done
