#!/bin/bash
# round 2, call 3k: the solve in two chunks (the last eighth covers the download of the first seven) -- tests, A/B of the tail fraction
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fold.py tests/test_gpu_drivers.py tests/test_gpu_solve_constraints.py tests/test_gpu_multi.py tests/test_gpu_matlab_glue.py -m gpu -x -q > gpurun_out/r3k_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r3k_sum.txt
tail -5 gpurun_out/r3k_pytest.log >> gpurun_out/r3k_sum.txt
for v in "OCS_B200_SOLVE_CHUNKS=0" "X=0" "OCS_B200_SOLVE_CHUNKS=6" "OCS_B200_SOLVE_CHUNKS=12" "OCS_B200_SOLVE_CHUNKS=4" "X=0" "OCS_B200_SOLVE_CHUNKS=0"; do
  env $v timeout 300 python bench.py --config cfg2solve --steps 3 --warmup 1 --no-cpu 2>/dev/null > gpurun_out/r3k_solve.json
  python -c "
import sys,json
j=json.load(open('gpurun_out/r3k_solve.json')); s=j['solve']; d=s['device_time']
print('$v', 'solve ms %.1f'%(j['ms_per_step']), 'per it', {k:round(x,3) for k,x in d['ms_per_iteration'].items()}, 'evals %.2f'%s['mean_evaluations'], 'launches', j.get('gpu_launches'), 'conv', s['converged'])
" >> gpurun_out/r3k_sum.txt 2>&1
done
cat gpurun_out/r3k_sum.txt
