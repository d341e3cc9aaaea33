#!/bin/bash
# round 2, call 3f: L2 re-read SPG update kernel as the default -- identity tests, unroll A/B; copy threads of the pageable path
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fold.py tests/test_gpu_drivers.py tests/test_gpu_solve_constraints.py -m gpu -x -q > gpurun_out/r3f_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r3f_sum.txt
tail -3 gpurun_out/r3f_pytest.log >> gpurun_out/r3f_sum.txt
for v in "X=0" "OCS_B200_SPG_UNROLL=2" "OCS_B200_SPG_UNROLL=8" "OCS_B200_SPG_L2=0" "X=0"; do
  env $v timeout 300 python bench.py --config cfg2solve --steps 1 --warmup 1 --no-cpu 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read()); s=j['solve']; d=s['device_time']
print('$v', 'solve ms %.1f'%(j['ms_per_step']), 'per it', {k:round(x,3) for k,x in d['ms_per_iteration'].items()}, 'update GB/s %.0f'%d['update_gbs'], 'evals %.2f'%s['mean_evaluations'])
" >> gpurun_out/r3f_sum.txt 2>&1
done
nproc >> gpurun_out/r3f_sum.txt; python -c "import os; print('cpus', os.cpu_count(), len(os.sched_getaffinity(0)))" >> gpurun_out/r3f_sum.txt
for v in "X=0" "OCS_B200_STAGER_THREADS=2" "OCS_B200_STAGER_THREADS=6" "OCS_B200_STAGER_THREADS=8" "OCS_B200_STAGER_THREADS=12" "OCS_B200_STAGER_SLOT_MB=8" "OCS_B200_STAGER_SLOT_MB=2" "X=0"; do
  env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-configs 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read()); e=j['e2e']
print('$v', 'e2e pinned ms %.2f pageable ms %.2f x%.2f'%(e['ms_per_step'], e['e2e_pageable']['ms_per_step'], e['e2e_pageable']['vs_pinned']))
" >> gpurun_out/r3f_sum.txt 2>&1
done
cat gpurun_out/r3f_sum.txt
