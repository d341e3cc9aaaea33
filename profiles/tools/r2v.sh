#!/bin/bash
# round 2, call v: full GPU test suite and the default bench line of the current build (+ the reference arm)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2v_sum.txt
tail -4 gpurun_out/r2v_pytest.log >> gpurun_out/r2v_sum.txt
timeout 1200 python bench.py > gpurun_out/r2v_bench_n1.json 2> gpurun_out/r2v_bench_n1.err; echo "bench rc=$?" >> gpurun_out/r2v_sum.txt
timeout 300 python bench.py --impl reference > gpurun_out/r2v_bench_ref.json 2> gpurun_out/r2v_bench_ref.err; echo "ref rc=$?" >> gpurun_out/r2v_sum.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> gpurun_out/r2v_sum.txt 2>&1
cat gpurun_out/r2v_sum.txt
