#!/bin/bash
# round 2, call 3c: step sizes as constant-bank operands (uniform grid): A/B on the built-in kernels
mkdir -p gpurun_out
bash profiles/tools/ab.sh "default unih default unih" "cfg2 cfg2v cfg2p" > gpurun_out/r3c_sum.txt 2>&1
cat gpurun_out/r3c_sum.txt
