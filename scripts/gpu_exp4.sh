set -x
mkdir -p gpurun_out
timeout 200 python scripts/dbg1.py > gpurun_out/dbg1.log 2>&1
echo done
