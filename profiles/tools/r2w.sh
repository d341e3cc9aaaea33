#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --config cfg2solve --no-cpu > gpurun_out/r2w_solve.json 2> gpurun_out/r2w_solve.err
python -c "import json; d=json.load(open('gpurun_out/r2w_solve.json')); print(d['ms_per_step'], json.dumps(d['solve'], indent=0))"
