#!/bin/bash
# round 2, call 3p (2 GPUs): the final build under torchrun at 2 ranks -- bench line and reference arm as the driver launches them
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r3p_bench_n2.json 2> gpurun_out/r3p_bench_n2.err; echo "bench n2 rc=$?" > gpurun_out/r3p_sum.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/r3p_ref_n2.json 2> gpurun_out/r3p_ref_n2.err; echo "ref n2 rc=$?" >> gpurun_out/r3p_sum.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r3p_pytest.log 2>&1; echo "pytest multi rc=$?" >> gpurun_out/r3p_sum.txt
tail -2 gpurun_out/r3p_pytest.log >> gpurun_out/r3p_sum.txt
python - <<'PY' >> gpurun_out/r3p_sum.txt
import json
d=json.load(open('gpurun_out/r3p_bench_n2.json'))
print('value %.4e ms %.3f'%(d['value'], d['ms_per_step']), d.get('per_rank',{}).get('ms_per_step'), d.get('gather_check'))
print('e2e %.3e'%d['e2e']['value'], d['e2e'].get('frac_of_pcie_ceiling'), 'solve %.3e %.1f ms'%(d['solve']['value'], d['solve']['ms_per_solve_batch']), 'multi %.3e'%d['c_abi_multi']['value'])
PY
cat gpurun_out/r3p_sum.txt
