#!/bin/bash
# round 2, call o: lanes in fb_sweep (late convergence check), bench lines of cfg3 / cfg3fb with the step view
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_lanes.py tests/test_gpu_parity.py tests/test_codegen.py -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2o_sum.txt
tail -5 gpurun_out/r2o_pytest.log >> gpurun_out/r2o_sum.txt
timeout 600 python bench.py --config cfg3 --no-cpu > gpurun_out/r2o_bench_cfg3.json 2> gpurun_out/r2o_bench_cfg3.err; echo "cfg3 rc=$?" >> gpurun_out/r2o_sum.txt
timeout 600 python bench.py --config cfg3fb --no-cpu > gpurun_out/r2o_bench_cfg3fb.json 2> gpurun_out/r2o_bench_cfg3fb.err; echo "cfg3fb rc=$?" >> gpurun_out/r2o_sum.txt
OCS_B200_LANES=0 timeout 600 python bench.py --config cfg3fb --no-cpu > gpurun_out/r2o_bench_cfg3fb_nolanes.json 2> gpurun_out/r2o_bench_cfg3fb_nolanes.err; echo "cfg3fb nolanes rc=$?" >> gpurun_out/r2o_sum.txt
timeout 600 python bench.py --config fbsweep --no-cpu > gpurun_out/r2o_bench_fbsweep.json 2> gpurun_out/r2o_bench_fbsweep.err; echo "fbsweep rc=$?" >> gpurun_out/r2o_sum.txt
python - <<'PY' >> gpurun_out/r2o_sum.txt
import json
for n in ("cfg3","cfg3fb","cfg3fb_nolanes","fbsweep"):
    try:
        j=json.load(open("gpurun_out/r2o_bench_%s.json"%n)); r=j["roofline"]
        print(n, "%.3e"%j["value"], "ms/step %.3f"%j["ms_per_step"], r["bound"], "frac %.3f"%r["frac"], {k:r.get(k) for k in ("step_hbm_frac","step_fp64_frac","step")}, [(k["kernel"][4:],round(k["ms"],3)) for k in r["kernels"]], "e2e", j.get("e2e",{}) and j["e2e"].get("value"))
    except Exception as e: print(n,"ERR",e)
PY
cat gpurun_out/r2o_sum.txt
