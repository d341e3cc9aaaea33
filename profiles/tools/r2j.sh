#!/bin/bash
# round 2, call j: resident warps per SM of the generated-problem sweeps (100k candidates = 21.1 warps per SM: 8 resident
# warps take 3 rounds, 12 take 2) and block granularity of the last round, after the FMA-shaped right-hand sides
mkdir -p gpurun_out
E1="OCS_LB_REV_GRID=3;OCS_LB_FWD_GRID=3;OCS_LB_FWD_DENSE=3"
E2="OCS_THREADS=64;OCS_TILE=16;OCS_LB_REV_GRID=6;OCS_LB_FWD_GRID=6;OCS_MINBLOCKS_FB=6;OCS_LB_FWD_DENSE=6"
E3="OCS_THREADS=32;OCS_TILE=8;OCS_LB_REV_GRID=12;OCS_LB_FWD_GRID=12;OCS_MINBLOCKS_FB=12;OCS_LB_FWD_DENSE=12"
E4="OCS_THREADS=64;OCS_LB_FWD_DENSE=4;OCS_LB_REV_GRID=4;OCS_LB_FWD_GRID=4;OCS_MINBLOCKS_FB=4"
E5="OCS_LB_FWD_GRID=2;OCS_LB_FWD_DENSE=2;OCS_LB_REV_GRID=2;OCS_MINBLOCKS_FB=2"
E6="OCS_THREADS=96;OCS_TILE=16;OCS_LB_REV_GRID=4;OCS_LB_FWD_GRID=4;OCS_MINBLOCKS_FB=4;OCS_LB_FWD_DENSE=4"
bash profiles/tools/ab2.sh "none $E1 $E2 $E3 $E4 $E5 $E6" "cfg3 cfg3fb" > gpurun_out/r2j_sum.txt 2>&1
cat gpurun_out/r2j_sum.txt
