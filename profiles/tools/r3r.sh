#!/bin/bash
# round 2, call 3r (2 GPUs): gather to rank 0 by peer copies over CUDA IPC (bench.py --gather peer, the default) against NCCL send/recv
mkdir -p gpurun_out
for g in peer root peer; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 3 --gather $g --no-e2e --no-configs --no-cpu > gpurun_out/r3r_bench_n2_$g.json 2> gpurun_out/r3r_bench_n2_$g.err; echo "bench n2 $g rc=$?" >> gpurun_out/r3r_sum.txt
python - $g <<'PY' >> gpurun_out/r3r_sum.txt
import json,sys
d=json.load(open('gpurun_out/r3r_bench_n2_%s.json'%sys.argv[1]))
print(sys.argv[1], 'value %.4e ms %.3f'%(d['value'], d['ms_per_step']), d.get('per_rank',{}).get('ms_per_step'), d.get('gather_check'))
PY
done
tail -5 gpurun_out/r3r_bench_n2_peer.err >> gpurun_out/r3r_sum.txt
cat gpurun_out/r3r_sum.txt
