# round 2, GPU call K: the record — ncu launch list and full capture of the bench command at config 3, smoke(), default bench.py
set -x
mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --repeats 1 --no-e2e --no-ref-gpu --cpu-sample 0 --c5-iters 0"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_c3_r2.csv $BENCH > gpurun_out/ncu_launch_r2k.log 2>&1
timeout 700 ncu --set full --clock-control none --import-source on -k regex:"estep_tc_kernel|mstep_tc_kernel" -s 6 -c 2 -f -o gpurun_out/prof_c3_r2 $BENCH > gpurun_out/ncu_full_r2k.log 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2k.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r2k.log
timeout 900 python bench.py > gpurun_out/bench_default_r2k.json 2> gpurun_out/bench_default_r2k.err; echo "bench rc=$?" >> gpurun_out/bench_default_r2k.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_r2k.json 2> gpurun_out/bench_reference_r2k.err; echo "bench rc=$?" >> gpurun_out/bench_reference_r2k.err
echo done
