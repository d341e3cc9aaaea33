#!/bin/bash
# round 2, call y: resident warps of the generated problem's forward sweep over the control grid (fb_sweep): it is the one
# latency-bound kernel of cfg3fb (long_scoreboard 3.7 warps per issue at 16 resident warps) and needs only 80 registers
mkdir -p gpurun_out
bash profiles/tools/ab2.sh "none OCS_LB_FWD_GRID=10 OCS_LB_FWD_GRID=12 OCS_LB_FWD_GRID=16 none" "cfg3fb" > gpurun_out/r2y_sum.txt 2>&1
cat gpurun_out/r2y_sum.txt
