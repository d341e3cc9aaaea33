#!/bin/bash
# round 2, call 3n: constant-operand projection kernel with double-buffered loads -- A/B on cfg3
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_codegen.py tests/test_gpu_golden.py tests/test_gpu_solve_constraints.py -m gpu -x -q > gpurun_out/r3n_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r3n_sum.txt
tail -3 gpurun_out/r3n_pytest.log >> gpurun_out/r3n_sum.txt
for v in "OCS_B200_PROJECT=lds" "X=0" "OCS_B200_PROJECT=lds" "X=0"; do
  env $v timeout 300 python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read()); ks=j['roofline']['kernels']; print('cfg3 $v','%.3e'%j['value'], 'ms/step %.3f'%j['ms_per_step'], ' '.join('%s=%.3f'%(k['kernel'][-22:],k['ms']) for k in ks))
" >> gpurun_out/r3n_sum.txt 2>&1
done
cat gpurun_out/r3n_sum.txt
