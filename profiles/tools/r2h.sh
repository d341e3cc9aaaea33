#!/bin/bash
# round 2, call h: GPU tests, the default bench line (with the configs block), the reference arm, and the ncu evidence of
# the same command: launch list (gpu__time_duration) and --set full captures of the cfg2 sweep kernels
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2h_sum.txt
tail -6 gpurun_out/r2h_pytest.log >> gpurun_out/r2h_sum.txt
timeout 900 python bench.py > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; echo "bench rc=$?" >> gpurun_out/r2h_sum.txt
timeout 300 python bench.py --impl reference > gpurun_out/r2h_bench_ref.json 2> gpurun_out/r2h_bench_ref.err; echo "ref rc=$?" >> gpurun_out/r2h_sum.txt
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-configs"
$CMD > gpurun_out/r2h_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2h_launches_cfg2.csv $CMD > gpurun_out/r2h_ncu1.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/r2h_sum.txt
ncu --set full --clock-control none --import-source on -k regex:'ocs_(fwd|rev)_test_fast' -s 6 -c 2 -o gpurun_out/r2h_cfg2 $CMD > gpurun_out/r2h_ncu2.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/r2h_sum.txt
CMD2="python bench.py --config cfg2p --steps 2 --warmup 3 --no-cpu --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:'ocs_(fwd|rev)_test_fast' -s 6 -c 2 -o gpurun_out/r2h_cfg2p $CMD2 > gpurun_out/r2h_ncu3.log 2>&1
echo "ncu cfg2p rc=$?" >> gpurun_out/r2h_sum.txt
CMD3="python bench.py --config cfg3 --steps 2 --warmup 3 --no-cpu --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:'ocs_(fwd|rev)_gen|dense_project' -s 6 -c 3 -o gpurun_out/r2h_cfg3 $CMD3 > gpurun_out/r2h_ncu4.log 2>&1
echo "ncu cfg3 rc=$?" >> gpurun_out/r2h_sum.txt
cat gpurun_out/r2h_sum.txt
