#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lanes.py tests/test_codegen.py -m gpu -x -q > gpurun_out/r3b_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r3b_sum.txt
tail -4 gpurun_out/r3b_pytest.log >> gpurun_out/r3b_sum.txt
timeout 300 python profiles/tools/_fb_overhead.py 2>&1 | tail -7 >> gpurun_out/r3b_sum.txt
timeout 600 python bench.py --config cfg3fb --no-cpu > gpurun_out/r3b_cfg3fb.json 2> gpurun_out/r3b_cfg3fb.err
python -c "
import json; d=json.load(open('gpurun_out/r3b_cfg3fb.json')); print('cfg3fb value %.3e e2e %.3e'%(d['value'], d['e2e']['value']))" >> gpurun_out/r3b_sum.txt
cat gpurun_out/r3b_sum.txt
