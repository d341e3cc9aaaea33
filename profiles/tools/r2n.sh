#!/bin/bash
# round 2, call n: two lanes (half batches on two streams) -- tests, then cfg3 / cfg3fb / cfg2 with the automatic decision and
# with lanes forbidden (OCS_B200_LANES=0)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2n_sum.txt
tail -5 gpurun_out/r2n_pytest.log >> gpurun_out/r2n_sum.txt
for lanes in auto 0; do
  if [ $lanes = auto ]; then unset OCS_B200_LANES; else export OCS_B200_LANES=$lanes; fi
  echo "== OCS_B200_LANES=$lanes" >> gpurun_out/r2n_sum.txt
  bash profiles/tools/ab2.sh "none" "cfg3 cfg3fb cfg2" >> gpurun_out/r2n_sum.txt 2>&1
done
unset OCS_B200_LANES
timeout 600 python bench.py --config cfg3 --no-cpu > gpurun_out/r2n_bench_cfg3.json 2> gpurun_out/r2n_bench_cfg3.err; echo "cfg3 rc=$?" >> gpurun_out/r2n_sum.txt
timeout 600 python bench.py --config cfg3fb --no-cpu > gpurun_out/r2n_bench_cfg3fb.json 2> gpurun_out/r2n_bench_cfg3fb.err; echo "cfg3fb rc=$?" >> gpurun_out/r2n_sum.txt
cat gpurun_out/r2n_sum.txt
