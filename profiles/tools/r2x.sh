#!/bin/bash
# round 2, call x: (1) does the 8 MiB row stride of the coefficient arrays (Bt = 2^20) hurt the solver's update kernel?  the same
# solve at Bt = 2^20 + 1536 and 2^20 + 4608; (2) the whole GPU suite with lanes forced wherever a batch can be halved
mkdir -p gpurun_out
for bt in 1048576 1050112 1053184; do
  timeout 600 python bench.py --config cfg2solve --no-cpu --bt $bt > gpurun_out/r2x_solve_$bt.json 2> gpurun_out/r2x_solve_$bt.err
  python -c "
import json; d=json.load(open('gpurun_out/r2x_solve_$bt.json')); t=d['solve']['device_time']; print($bt, 'ms/solve %.1f'%d['ms_per_step'], 'per iteration', {k: round(v,3) for k,v in t['ms_per_iteration'].items()}, 'update GB/s %.0f'%t['update_gbs'])"
done > gpurun_out/r2x_sum.txt 2>&1
OCS_B200_LANES=1 timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_pytest_lanes1.log 2>&1; echo "pytest (OCS_B200_LANES=1) rc=$?" >> gpurun_out/r2x_sum.txt
tail -4 gpurun_out/r2x_pytest_lanes1.log >> gpurun_out/r2x_sum.txt
cat gpurun_out/r2x_sum.txt
