#!/bin/bash
# round 2, call 3h: pinned vOpt written by the device as trajectories stop (publish_minimisers_kernel) -- tests, A/B on cfg2solve
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fold.py tests/test_gpu_drivers.py tests/test_gpu_multi.py tests/test_gpu_host_path.py -m gpu -x -q > gpurun_out/r3h_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r3h_sum.txt
tail -5 gpurun_out/r3h_pytest.log >> gpurun_out/r3h_sum.txt
for v in "X=0" "OCS_B200_SOLVE_PUBLISH=0" "X=0" "OCS_B200_SOLVE_PUBLISH=0"; do
  env $v timeout 300 python bench.py --config cfg2solve --steps 3 --warmup 1 --no-cpu 2>/dev/null > gpurun_out/r3h_solve.json
  python -c "
import sys,json
j=json.load(open('gpurun_out/r3h_solve.json')); s=j['solve']; d=s['device_time']
print('$v', 'solve ms %.1f'%(j['ms_per_step']), 'per it', {k:round(x,3) for k,x in d['ms_per_iteration'].items()}, 'evals %.2f'%s['mean_evaluations'], 'launches', j.get('gpu_launches'))
" >> gpurun_out/r3h_sum.txt 2>&1
done
cp gpurun_out/r3h_solve.json gpurun_out/r3h_solve_nopublish.json
timeout 300 python bench.py --config cfg2solve --steps 3 --warmup 1 --no-cpu 2>/dev/null > gpurun_out/r3h_solve.json
cat gpurun_out/r3h_sum.txt
