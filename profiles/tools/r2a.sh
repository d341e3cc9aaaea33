#!/bin/bash
# round 2, call a: where do the generated-problem (cfg3) kernels stand -- block size / unroll / register-cap A/B
# and one ncu --set full capture of the forward, reverse and projection kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
bash profiles/tools/ab2.sh "none OCS_REV_UNROLL=1 OCS_THREADS=64 OCS_THREADS=32 OCS_THREADS=32;OCS_MINBLOCKS_REV=11;OCS_MINBLOCKS_FWD=11 OCS_THREADS=64;OCS_MINBLOCKS_REV=5;OCS_MINBLOCKS_FWD=5 OCS_THREADS=96;OCS_MINBLOCKS_REV=3;OCS_MINBLOCKS_FWD=3" "cfg3" > gpurun_out/r2a_ab.txt 2>&1
bash profiles/tools/ab2.sh "none OCS_THREADS=64 OCS_THREADS=32;OCS_MINBLOCKS_FB=11;OCS_MINBLOCKS_FWD=11" "cfg3fb" >> gpurun_out/r2a_ab.txt 2>&1
CMD="python bench.py --config cfg3 --steps 3 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/r2a_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'ocs_(fwd|rev)_gen|dense_project' -s 9 -c 3 -o gpurun_out/r2a_cfg3 $CMD > gpurun_out/r2a_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/r2a_ab.txt
cat gpurun_out/r2a_ab.txt
