#!/bin/bash
# round 2, call 3l: the final build -- whole GPU suite, smoke, default bench line (configs block), reference arm, and the ncu launch
# lists of the cfg2 evaluation and of the batched solve
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r3l_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r3l_sum.txt
tail -4 gpurun_out/r3l_pytest.log >> gpurun_out/r3l_sum.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> gpurun_out/r3l_sum.txt 2>&1
timeout 1200 python bench.py > gpurun_out/r3l_bench_n1.json 2> gpurun_out/r3l_bench_n1.err; echo "bench rc=$?" >> gpurun_out/r3l_sum.txt
timeout 300 python bench.py --impl reference > gpurun_out/r3l_bench_ref.json 2> gpurun_out/r3l_bench_ref.err; echo "ref rc=$?" >> gpurun_out/r3l_sum.txt
timeout 300 python bench.py --config cfg2solve --no-cpu > gpurun_out/r3l_bench_cfg2solve.json 2> gpurun_out/r3l_bench_cfg2solve.err; echo "solve rc=$?" >> gpurun_out/r3l_sum.txt
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-configs"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3l_launches_cfg2.csv $CMD > gpurun_out/r3l_ncu1.log 2>&1
echo "ncu launches cfg2 rc=$?" >> gpurun_out/r3l_sum.txt
CMD2="python bench.py --config cfg2solve --steps 1 --warmup 1 --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3l_launches_cfg2solve.csv $CMD2 > gpurun_out/r3l_ncu2.log 2>&1
echo "ncu launches cfg2solve rc=$?" >> gpurun_out/r3l_sum.txt
python - <<'PY' >> gpurun_out/r3l_sum.txt
import json
d=json.load(open('gpurun_out/r3l_bench_n1.json'))
print('value %.4e ms %.3f frac %.3f'%(d['value'], d['ms_per_step'], d['roofline']['frac']), [(k['kernel'][4:22], round(k['ms'],3)) for k in d['roofline']['kernels']])
print('e2e %.3e pageable x%.2f'%(d['e2e']['value'], d['e2e']['e2e_pageable']['vs_pinned']), 'solve %.1f ms'%(d['solve']['ms_per_solve_batch']))
for k,c in d['configs'].items(): print(k, '%.3e'%c['value'], c.get('frac'), [(q['kernel'][4:22], round(q['ms'],3)) for q in c.get('kernels',[])])
r=json.load(open('gpurun_out/r3l_bench_ref.json')); print('reference arm %.3e'%r['value'], r['cpu_baseline']['cores'])
PY
cat gpurun_out/r3l_sum.txt
