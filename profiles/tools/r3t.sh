#!/bin/bash
# round 2, call 3t (8 GPUs): peer-copy gather with a 4-byte token gathered to the root as completion, against the NCCL gather (same box)
mkdir -p gpurun_out; rm -f gpurun_out/r3t_sum.txt
for g in peer root; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 20 --warmup 3 --gather $g --no-e2e --no-configs --no-cpu > gpurun_out/r3t_bench_n8_$g.json 2> gpurun_out/r3t_bench_n8_$g.err; echo "bench n8 $g rc=$?" >> gpurun_out/r3t_sum.txt
python - $g <<'PY' >> gpurun_out/r3t_sum.txt
import json,sys
d=json.load(open('gpurun_out/r3t_bench_n8_%s.json'%sys.argv[1]))
print(sys.argv[1], 'value %.4e ms %.3f'%(d['value'], d['ms_per_step']), [round(x,3) for x in d.get('per_rank',{}).get('ms_per_step',[])], d.get('gather_check'))
PY
done
cat gpurun_out/r3t_sum.txt
