#!/bin/bash
# round 2, call 3q (8 GPUs, final build): the multi-GPU tests of the one-handle C ABI path with real shards, then the bench under torchrun
# at 8 ranks (NCCL gather of J, x(T) to rank 0 inside the timed region; e2e; solve; the one-process multi-device handle)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r3q_pytest.log 2>&1; echo "pytest multi rc=$?" > gpurun_out/r3q_sum.txt
tail -3 gpurun_out/r3q_pytest.log >> gpurun_out/r3q_sum.txt
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r3q_bench_n8.json 2> gpurun_out/r3q_bench_n8.err; echo "bench n8 rc=$?" >> gpurun_out/r3q_sum.txt
grep -m3 -E "NVLS|Connected all|nranks" gpurun_out/r3q_bench_n8.err >> gpurun_out/r3q_sum.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --impl reference --steps 3 --warmup 1 > gpurun_out/r3q_ref_n8.json 2> gpurun_out/r3q_ref_n8.err; echo "ref n8 rc=$?" >> gpurun_out/r3q_sum.txt
cat gpurun_out/r3q_sum.txt
