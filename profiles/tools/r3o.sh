#!/bin/bash
# round 2, call 3o: constant-operand projection, one control per thread at 6 blocks/SM (c1) against two per thread at 4 and the LDS kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_codegen.py tests/test_gpu_solve_constraints.py -m gpu -x -q > gpurun_out/r3o_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r3o_sum.txt
OCS_B200_PROJECT=c1 timeout 600 python -m pytest tests/test_codegen.py tests/test_gpu_solve_constraints.py -m gpu -x -q >> gpurun_out/r3o_pytest.log 2>&1; echo "pytest c1 rc=$?" >> gpurun_out/r3o_sum.txt
for v in "OCS_B200_PROJECT=lds" "X=0" "OCS_B200_PROJECT=c1" "OCS_B200_PROJECT=lds" "X=0" "OCS_B200_PROJECT=c1"; do
  env $v timeout 300 python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read()); ks=j['roofline']['kernels']; print('cfg3 $v','%.3e'%j['value'], 'ms/step %.3f'%j['ms_per_step'], ' '.join('%s=%.3f'%(k['kernel'][-22:],k['ms']) for k in ks))
" >> gpurun_out/r3o_sum.txt 2>&1
done
cat gpurun_out/r3o_sum.txt
