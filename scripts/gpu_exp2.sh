set -x
mkdir -p gpurun_out
timeout 400 python scripts/exp_gsplit.py > gpurun_out/exp_gsplit.json 2> gpurun_out/exp_gsplit.err; echo "rc=$?" >> gpurun_out/exp_gsplit.err
echo done
