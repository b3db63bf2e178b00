# round 2, GPU call S: compute-sanitizer over the device-side finalisation (memcheck of the whole small run, racecheck and synccheck of finalize_params_kernel)
set -x
mkdir -p gpurun_out
SAN_N=6000 timeout 300 compute-sanitizer --tool memcheck python scripts/san_run.py > gpurun_out/san_mem_r2s.log 2>&1
SAN_N=3000 timeout 300 compute-sanitizer --tool racecheck --kernel-name kernel_substring=finalize_params python scripts/san_run.py > gpurun_out/san_race_r2s.log 2>&1
SAN_N=3000 SAN_D=16 SAN_K=20 timeout 300 compute-sanitizer --tool racecheck --kernel-name kernel_substring=finalize_params python scripts/san_run.py > gpurun_out/san_race16_r2s.log 2>&1
SAN_N=3000 timeout 300 compute-sanitizer --tool synccheck --kernel-name kernel_substring=finalize_params python scripts/san_run.py > gpurun_out/san_sync_r2s.log 2>&1
echo done
