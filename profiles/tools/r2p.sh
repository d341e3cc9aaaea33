#!/bin/bash
# round 2, call p: 64-thread blocks for the generated-problem sweeps with lanes on (value = whole step with lanes; kernel ms = alone)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lanes.py -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2p_sum.txt
tail -3 gpurun_out/r2p_pytest.log >> gpurun_out/r2p_sum.txt
B="OCS_THREADS=64;OCS_LB_FWD_DENSE=4;OCS_LB_REV_GRID=4;OCS_LB_FWD_GRID=4;OCS_MINBLOCKS_FB=4"
C="OCS_THREADS=64;OCS_LB_FWD_DENSE=6;OCS_LB_REV_GRID=4;OCS_LB_FWD_GRID=8;OCS_MINBLOCKS_FB=4"
D="OCS_THREADS=64;OCS_TILE=16;OCS_LB_FWD_DENSE=6;OCS_LB_REV_GRID=4;OCS_LB_FWD_GRID=8;OCS_MINBLOCKS_FB=4"
bash profiles/tools/ab2.sh "none $B $C $D none" "cfg3 cfg3fb" >> gpurun_out/r2p_sum.txt 2>&1
cat gpurun_out/r2p_sum.txt
