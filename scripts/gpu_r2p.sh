# round 2, GPU call P: the record after the device-side finalisation — ncu launch list and full capture (E-step, M-step, finalize_params_kernel)
set -x
mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --repeats 1 --no-e2e --no-ref-gpu --cpu-sample 0 --c5-iters 0"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_c3_r2b.csv $BENCH > gpurun_out/ncu_launch_r2p.log 2>&1
timeout 700 ncu --set full --clock-control none --import-source on -k regex:"estep_tc_kernel|mstep_tc_kernel|finalize_params_kernel" -s 9 -c 3 -f -o gpurun_out/prof_c3_r2b $BENCH > gpurun_out/ncu_full_r2p.log 2>&1
echo done
