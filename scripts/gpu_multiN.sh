# usage (from the repo root):  gpurun --gpus N -- 'N=<N> bash scripts/gpu_multiN.sh'
set -x
N=${N:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_c3_g${N}.json 2> gpurun_out/bench_c3_g${N}.err
if [ "$N" = "2" ]; then timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "two_gpus" > gpurun_out/pytest_multi2.log 2>&1; fi
echo done
