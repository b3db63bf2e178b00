set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus2.txt
timeout 600 python -m pytest tests -m gpu -q -k "two_gpus or cli" > gpurun_out/pytest_multi2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_c3_g2.json 2> gpurun_out/bench_c3_g2.err
echo done
