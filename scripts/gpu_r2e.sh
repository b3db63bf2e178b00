# round 2, GPU call E: rotation-symmetric M-step builder (one instruction stream), two-stage E-step epilogue (squares -> TMEM -> log-sum-exp/stores)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "estep_parity or mstep_constants_parity or tensor_path_large or tight" > gpurun_out/pytest_r2e_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2e_quick.log
export GMM_EXP_N=4000000
timeout 120 python scripts/prof_run.py > gpurun_out/time_r2e.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_r2e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2e.log
timeout 600 python scripts/exp_acc.py all default > gpurun_out/exp_acc_r2e.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --repeats 3 --c5-iters 0 --no-ref-gpu --cpu-sample 0 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "bench rc=$?" >> gpurun_out/bench_r2e.err
echo done
