#!/bin/bash
# round 2, call k: ncu --set full of the generated-problem sweeps after the FMA-shaped right-hand sides (cfg3, cfg3fb)
mkdir -p gpurun_out
CMD3="python bench.py --config cfg3 --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD3 > gpurun_out/r2k_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'ocs_(fwd|rev)_gen|dense_project' -s 6 -c 3 -o gpurun_out/r2k_cfg3 $CMD3 > gpurun_out/r2k_ncu3.log 2>&1
echo "ncu cfg3 rc=$?" > gpurun_out/r2k_sum.txt
CMD4="python bench.py --config cfg3fb --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD4 > gpurun_out/r2k_plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'ocs_(fwd|fbrev)_gen' -s 8 -c 2 -o gpurun_out/r2k_cfg3fb $CMD4 > gpurun_out/r2k_ncu4.log 2>&1
echo "ncu cfg3fb rc=$?" >> gpurun_out/r2k_sum.txt
cat gpurun_out/r2k_sum.txt
