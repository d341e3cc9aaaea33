#!/bin/bash
# round 2, call q: the build with FMA-shaped right-hand sides, two lanes and 64-thread blocks for several-state problems:
# GPU tests, the default bench line (configs block), ncu launch list + --set full captures of cfg3 and cfg3fb
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2q_sum.txt
tail -4 gpurun_out/r2q_pytest.log >> gpurun_out/r2q_sum.txt
timeout 900 python bench.py > gpurun_out/r2q_bench_n1.json 2> gpurun_out/r2q_bench_n1.err; echo "bench rc=$?" >> gpurun_out/r2q_sum.txt
CMD3="python bench.py --config cfg3 --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD3 > gpurun_out/r2q_plain3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2q_launches_cfg3.csv $CMD3 > gpurun_out/r2q_ncu3a.log 2>&1
echo "ncu launches cfg3 rc=$?" >> gpurun_out/r2q_sum.txt
OCS_B200_LANES=0 ncu --set full --clock-control none --import-source on -k regex:'ocs_(fwd|rev)_gen|dense_project' -s 6 -c 3 -o gpurun_out/r2q_cfg3 $CMD3 > gpurun_out/r2q_ncu3.log 2>&1
echo "ncu cfg3 rc=$?" >> gpurun_out/r2q_sum.txt
CMD4="python bench.py --config cfg3fb --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD4 > gpurun_out/r2q_plain4.log 2>&1 && \
OCS_B200_LANES=0 ncu --set full --clock-control none --import-source on -k regex:'ocs_(fwd|fbrev)_gen' -s 8 -c 2 -o gpurun_out/r2q_cfg3fb $CMD4 > gpurun_out/r2q_ncu4.log 2>&1
echo "ncu cfg3fb rc=$?" >> gpurun_out/r2q_sum.txt
cat gpurun_out/r2q_sum.txt
