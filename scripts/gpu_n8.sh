# round 2, the one 8-GPU call: peer-memory all-reduce and sharded C-ABI parity at G = 8, then bench.py under torchrun at N = 8
# (short: the call is charged 8x).   gpurun --gpus 8 -- 'bash scripts/gpu_n8.sh'
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_g8.txt 2>&1
timeout 240 python -m pytest -m gpu -q "tests/test_gpu_scale.py::test_peer_memory_allreduce_equals_nccl[8]" "tests/test_gpu_scale.py::test_sharded_c_abi_equals_single_gpu_and_oracle[auto-8]" "tests/test_gpu_scale.py::test_sharded_c_abi_equals_single_gpu_and_oracle[simt-8]" > gpurun_out/pytest_multi_g8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi_g8.log
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --steps 20 --warmup 3 --no-ref-gpu --cpu-sample 0 --c5-iters 0 > gpurun_out/bench_r2_g8.json 2> gpurun_out/bench_r2_g8.err; echo "bench rc=$?" >> gpurun_out/bench_r2_g8.err
echo done
