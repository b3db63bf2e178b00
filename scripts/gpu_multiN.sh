set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_c3_g4_c.json 2> gpurun_out/bench_c3_g4_c.err
echo done
