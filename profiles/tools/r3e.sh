#!/bin/bash
# round 2, call 3e: SPG update kernel that re-reads the accepted rows from L2 (OCS_B200_SPG_L2) -- identity test, A/B on cfg2solve
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fold.py tests/test_gpu_drivers.py -m gpu -x -q > gpurun_out/r3e_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r3e_sum.txt
tail -3 gpurun_out/r3e_pytest.log >> gpurun_out/r3e_sum.txt
for v in 0 1 2 4 11 12 14 0 2; do
  OCS_B200_SPG_L2=$v timeout 300 python bench.py --config cfg2solve --steps 1 --warmup 1 --no-cpu 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read()); s=j['solve']; d=s['device_time']
print('L2=$v', 'solve ms %.1f'%(j['ms_per_step']), 'per it', {k:round(x,3) for k,x in d['ms_per_iteration'].items()}, 'update GB/s %.0f'%d['update_gbs'], 'evals %.2f'%s['mean_evaluations'])
" >> gpurun_out/r3e_sum.txt 2>&1
done
cat gpurun_out/r3e_sum.txt
