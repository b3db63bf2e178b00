set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus8.txt
for n in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_c3_g$n.json 2> gpurun_out/bench_c3_g$n.err
done
echo done
