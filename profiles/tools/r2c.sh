#!/bin/bash
# round 2, call c: all GPU tests (new: multi-handle, pageable ring, resident checks), cfg3 split timings, ncu of cfg3
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2c_ab.txt
tail -15 gpurun_out/r2c_pytest.log >> gpurun_out/r2c_ab.txt
bash profiles/tools/ab2.sh "none OCS_MINBLOCKS_REV=2;OCS_MINBLOCKS_FWD=2" "cfg3" >> gpurun_out/r2c_ab.txt 2>&1
CMD="python bench.py --config cfg3 --steps 3 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/r2c_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'ocs_(fwd|rev)_gen|dense_project' -s 9 -c 3 -o gpurun_out/r2c_cfg3 $CMD > gpurun_out/r2c_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/r2c_ab.txt
cat gpurun_out/r2c_ab.txt
