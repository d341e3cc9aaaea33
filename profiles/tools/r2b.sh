#!/bin/bash
# round 2, call b: parity of the reworked kernels (GPU tests), then A/B of async tile staging on the built-in kernels
# and the new generated-problem defaults (ring prefetch, 3 blocks/SM, new projection kernel).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2b_ab.txt
tail -5 gpurun_out/r2b_pytest.log >> gpurun_out/r2b_ab.txt
bash profiles/tools/ab.sh "default noasync" "cfg2 cfg2v cfg4" >> gpurun_out/r2b_ab.txt 2>&1
bash profiles/tools/ab2.sh "none OCS_RING=8 OCS_MINBLOCKS_REV=2;OCS_MINBLOCKS_FWD=2 OCS_MINBLOCKS_REV=4;OCS_MINBLOCKS_FWD=4 OCS_ASYNC_STAGE=0 OCS_RING=0" "cfg3" >> gpurun_out/r2b_ab.txt 2>&1
bash profiles/tools/ab2.sh "none OCS_FB_RING=0 OCS_FB_RING=8 OCS_MINBLOCKS_FB=2 OCS_MINBLOCKS_FB=4" "cfg3fb" >> gpurun_out/r2b_ab.txt 2>&1
cat gpurun_out/r2b_ab.txt
