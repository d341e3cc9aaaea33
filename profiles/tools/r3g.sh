#!/bin/bash
# round 2, call 3g: vOpt download under the last evaluation; evaluation-count histogram; ncu --set full of the cfg2 sweeps
# (constant step-size operands) and of the L2 re-read update kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fold.py tests/test_gpu_drivers.py tests/test_gpu_solve_constraints.py tests/test_gpu_host_path.py -m gpu -x -q > gpurun_out/r3g_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r3g_sum.txt
tail -3 gpurun_out/r3g_pytest.log >> gpurun_out/r3g_sum.txt
for v in "X=0" "OCS_B200_SPG_UNROLL=1" "X=0"; do
  env $v timeout 300 python bench.py --config cfg2solve --steps 3 --warmup 1 --no-cpu 2>/dev/null > gpurun_out/r3g_solve.json
  python -c "
import sys,json
j=json.load(open('gpurun_out/r3g_solve.json')); s=j['solve']; d=s['device_time']
print('$v', 'solve ms %.1f'%(j['ms_per_step']), 'per it', {k:round(x,3) for k,x in d['ms_per_iteration'].items()}, 'update GB/s %.0f'%d['update_gbs'], 'evals %.2f'%s['mean_evaluations'], d['evaluations_histogram'])
" >> gpurun_out/r3g_sum.txt 2>&1
done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-configs"
ncu --set full --clock-control none --import-source on -k regex:'ocs_(fwd|rev)_test_fast' -s 6 -c 2 -o gpurun_out/r3g_cfg2 $CMD > gpurun_out/r3g_ncu2.log 2>&1
echo "ncu cfg2 rc=$?" >> gpurun_out/r3g_sum.txt
CMD2="python bench.py --config cfg2solve --steps 1 --warmup 1 --no-cpu"
ncu --set full --clock-control none --import-source on -k regex:'spg_update|ocs_(fwd|rev)_test_fast_u1' -s 12 -c 3 -o gpurun_out/r3g_solve $CMD2 > gpurun_out/r3g_ncu3.log 2>&1
echo "ncu solve rc=$?" >> gpurun_out/r3g_sum.txt
cat gpurun_out/r3g_sum.txt
