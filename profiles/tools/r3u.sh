#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3u_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r3u_sum.txt
tail -3 gpurun_out/r3u_pytest.log >> gpurun_out/r3u_sum.txt
timeout 600 python bench.py > gpurun_out/r3u_bench_n1.json 2> gpurun_out/r3u_bench_n1.err; echo "bench rc=$?" >> gpurun_out/r3u_sum.txt
python - <<'PY' >> gpurun_out/r3u_sum.txt
import json
d=json.load(open('gpurun_out/r3u_bench_n1.json'))
print('value %.4e ms %.3f frac %.3f'%(d['value'], d['ms_per_step'], d['roofline']['frac']), [(k['kernel'][4:22], round(k['ms'],3)) for k in d['roofline']['kernels']])
print('e2e %.3e pageable x%.2f'%(d['e2e']['value'], d['e2e']['e2e_pageable']['vs_pinned']), 'solve %.1f ms'%(d['solve']['ms_per_solve_batch']))
for k,c in d['configs'].items(): print(k, '%.3e'%c['value'], c.get('frac'), [(q['kernel'][-22:], round(q['ms'],3)) for q in c.get('kernels',[])])
PY
cat gpurun_out/r3u_sum.txt
