# round 2, GPU call R: finalize_params_kernel with a thread per matrix element — parity, replay, phase stamps, bench
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_scale.py -m gpu -q -x -k "device_finalisation or em_iterations" > gpurun_out/pytest_r2r.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2r.log
GMM_EXP_N=1000000 GMM_EXP_ITERS=2 GMM_B200_LIB=cuda-gmm-mpi_b200/variants/libgmm_b200_finprof.so timeout 120 python scripts/prof_run.py > gpurun_out/finprof.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --repeats 3 --no-ref-gpu --cpu-sample 0 --c5-iters 0 > gpurun_out/bench_r2r.json 2> gpurun_out/bench_r2r.err
echo done
