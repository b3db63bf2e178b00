set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_c3_g8_c.json 2> gpurun_out/bench_c3_g8_c.err
echo done
