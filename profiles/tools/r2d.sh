#!/bin/bash
# round 2, call d (2 GPUs): all GPU tests with two devices visible (real shards in the multi-handle tests), the bench
# line at N=1 with the new blocks, and the torchrun N=2 line (gather to rank 0, per-rank times, c_abi_multi)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2d_sum.txt
tail -12 gpurun_out/r2d_pytest.log >> gpurun_out/r2d_sum.txt
timeout 900 python bench.py > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; echo "bench n1 rc=$?" >> gpurun_out/r2d_sum.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err; echo "bench n2 rc=$?" >> gpurun_out/r2d_sum.txt
OCS_B200_GATHER=p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2d_bench_n2_p2p.json 2> gpurun_out/r2d_bench_n2_p2p.err; echo "bench n2 p2p rc=$?" >> gpurun_out/r2d_sum.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2d_ref.json 2>gpurun_out/r2d_ref.err; echo "ref rc=$?" >> gpurun_out/r2d_sum.txt
tail -3 gpurun_out/r2d_bench_n1.err gpurun_out/r2d_bench_n2.err >> gpurun_out/r2d_sum.txt
cat gpurun_out/r2d_sum.txt
