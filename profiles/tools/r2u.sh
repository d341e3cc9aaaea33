#!/bin/bash
# round 2, call u: node-state prefetch of the built-in reverse kernels (cp.async ring depth / register prefetch depth) on the
# variants whose top stall is long_scoreboard (per-trajectory parameters, per-trajectory coefficients); guard-row test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lanes.py -m gpu -x -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2u_sum.txt
tail -3 gpurun_out/r2u_pytest.log >> gpurun_out/r2u_sum.txt
bash profiles/tools/ab.sh "default ring4 ring2 pf3 pf4 default" "cfg2p cfg2v cfg2" >> gpurun_out/r2u_sum.txt 2>&1
cat gpurun_out/r2u_sum.txt
