#!/bin/bash
# round 2, call 3s (8 GPUs, final build, peer-copy gather): the multi-GPU tests of the one-handle C ABI path with real shards, then the bench under torchrun
# at 8 ranks (NCCL gather of J, x(T) to rank 0 inside the timed region; e2e; solve; the one-process multi-device handle)
mkdir -p gpurun_out
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r3s_bench_n8.json 2> gpurun_out/r3s_bench_n8.err; echo "bench n8 rc=$?" > gpurun_out/r3s_sum.txt
grep -m3 -E "NVLS|Connected all|nranks" gpurun_out/r3s_bench_n8.err >> gpurun_out/r3s_sum.txt
python - <<'PY' >> gpurun_out/r3s_sum.txt
import json
d=json.load(open('gpurun_out/r3s_bench_n8.json'))
print('value %.4e ms %.3f'%(d['value'], d['ms_per_step']), d.get('per_rank',{}).get('ms_per_step'), d.get('gather_check'))
print('e2e %.3e'%d['e2e']['value'], 'solve %.1f ms'%d['solve']['ms_per_solve_batch'], 'multi %.3e'%d['c_abi_multi']['value'])
PY
cat gpurun_out/r3s_sum.txt
