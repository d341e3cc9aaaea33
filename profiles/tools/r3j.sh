#!/bin/bash
# round 2, call 3j: occupancy re-check of the built-in sweeps after the constant step sizes freed registers
mkdir -p gpurun_out
cp profiles/tools/ab.sh /tmp/ab.sh; sed -i 's#optimal-control-solvers_b200/build/variants#optimal-control-solvers_b200/build/variants#' /tmp/ab.sh
bash profiles/tools/ab.sh "default rev6 revc5 fwdc6 default rev6 revc5 fwdc6" "cfg2 cfg2v" > gpurun_out/r3j_sum.txt 2>&1
cat gpurun_out/r3j_sum.txt
