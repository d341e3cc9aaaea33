#!/bin/bash
# round 2, call i: FMA-shaped generated right-hand sides (codegen/fma_forms.py): parity tests of the generated problems, then
# cfg3 / cfg3fb with block sizes that change the number of resident warps per SM (100k candidates = 21.1 warps per SM)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_codegen.py tests/test_gpu_fullsize.py tests/test_gpu_equilibrium.py -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2i_sum.txt
tail -4 gpurun_out/r2i_pytest.log >> gpurun_out/r2i_sum.txt
bash profiles/tools/ab2.sh "none OCS_THREADS=64 OCS_THREADS=352 OCS_THREADS=32;OCS_TILE=16 OCS_THREADS=352;OCS_RING=2;OCS_FB_RING=2" "cfg3 cfg3fb" >> gpurun_out/r2i_sum.txt 2>&1
cat gpurun_out/r2i_sum.txt
