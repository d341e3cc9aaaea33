#!/bin/bash
# round 2, call g: GPU tests (recurrence of time exponentials in the parameter-sweep kernels), cfg2p / cfg2pv / cfg4
# timings, the cfg4 end-to-end path with the horizon-aware chunk schedule
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2g_sum.txt
tail -12 gpurun_out/r2g_pytest.log >> gpurun_out/r2g_sum.txt
bash profiles/tools/ab.sh "default" "cfg2p cfg2pv cfg2 cfg2v" >> gpurun_out/r2g_sum.txt 2>&1
timeout 600 python bench.py --config cfg4 --no-cpu --steps 5 > gpurun_out/r2g_bench_cfg4.json 2> gpurun_out/r2g_bench_cfg4.err; echo "cfg4 rc=$?" >> gpurun_out/r2g_sum.txt
cat gpurun_out/r2g_sum.txt
