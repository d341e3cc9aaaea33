#!/bin/bash
# round 2, call 3i: the device-written vOpt slowed the forward sweeps it overlapped (r3h) -- fewer blocks, and under the reverse sweep instead
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fold.py -m gpu -x -q -k pinned > gpurun_out/r3i_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r3i_sum.txt
OCS_B200_PUBLISH_AFTER_FWD=1 OCS_B200_PUBLISH_BLOCKS=4 timeout 600 python -m pytest tests/test_gpu_fold.py -m gpu -x -q -k pinned >> gpurun_out/r3i_pytest.log 2>&1; echo "pytest (after fwd) rc=$?" >> gpurun_out/r3i_sum.txt
for v in "OCS_B200_SOLVE_PUBLISH=0" "OCS_B200_PUBLISH_BLOCKS=16" "OCS_B200_PUBLISH_BLOCKS=4" "OCS_B200_PUBLISH_BLOCKS=1" "OCS_B200_PUBLISH_AFTER_FWD=1" "OCS_B200_PUBLISH_AFTER_FWD=1 OCS_B200_PUBLISH_BLOCKS=16" "OCS_B200_PUBLISH_AFTER_FWD=1 OCS_B200_PUBLISH_BLOCKS=4" "OCS_B200_SOLVE_PUBLISH=0"; do
  env $v timeout 300 python bench.py --config cfg2solve --steps 3 --warmup 1 --no-cpu 2>/dev/null > gpurun_out/r3i_solve.json
  python -c "
import sys,json
j=json.load(open('gpurun_out/r3i_solve.json')); s=j['solve']; d=s['device_time']
print('$v', 'solve ms %.1f'%(j['ms_per_step']), 'per it', {k:round(x,3) for k,x in d['ms_per_iteration'].items()}, 'launches', j.get('gpu_launches'))
" >> gpurun_out/r3i_sum.txt 2>&1
done
cat gpurun_out/r3i_sum.txt
