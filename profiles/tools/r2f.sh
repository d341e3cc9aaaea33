#!/bin/bash
# round 2, call f: GPU tests; parameter-sweep launch configuration A/B (default = spill-free (1,5;1,4) vs (1,8;1,5));
# cfg4 with the spill-free tail kernel and the pipelined host path; ncu of the projection + cfg2p kernels
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2f_sum.txt
tail -12 gpurun_out/r2f_pytest.log >> gpurun_out/r2f_sum.txt
bash profiles/tools/ab.sh "default f18" "cfg2p cfg2pv cfg4" >> gpurun_out/r2f_sum.txt 2>&1
timeout 600 python bench.py --config cfg4 --no-cpu --steps 5 > gpurun_out/r2f_bench_cfg4.json 2> gpurun_out/r2f_bench_cfg4.err; echo "cfg4 rc=$?" >> gpurun_out/r2f_sum.txt
CMD="python bench.py --config cfg2p --steps 3 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/r2f_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'ocs_(fwd|rev)_test' -s 6 -c 2 -o gpurun_out/r2f_cfg2p $CMD > gpurun_out/r2f_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/r2f_sum.txt
cat gpurun_out/r2f_sum.txt
