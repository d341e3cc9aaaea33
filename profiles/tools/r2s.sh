#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_solve_constraints.py tests/test_gpu_drivers.py -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/r2s_pytest.log
