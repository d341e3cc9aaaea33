#!/bin/bash
# round 2, call e: GPU tests; per-trajectory-parameter launch configurations (cfg2p / cfg2pv); cfg3 block sizes with the
# new projection kernel; the N=1 bench line (pageable path with non-temporal copies)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2e_sum.txt
tail -12 gpurun_out/r2e_pytest.log >> gpurun_out/r2e_sum.txt
bash profiles/tools/ab.sh "default f16 f18 f24 f23" "cfg2p cfg2pv" >> gpurun_out/r2e_sum.txt 2>&1
bash profiles/tools/ab2.sh "none OCS_THREADS=64;OCS_MINBLOCKS_REV=6;OCS_MINBLOCKS_FWD=6 OCS_THREADS=32;OCS_MINBLOCKS_REV=12;OCS_MINBLOCKS_FWD=12 OCS_THREADS=96;OCS_MINBLOCKS_REV=4;OCS_MINBLOCKS_FWD=4" "cfg3" >> gpurun_out/r2e_sum.txt 2>&1
bash profiles/tools/ab2.sh "none OCS_THREADS=64;OCS_MINBLOCKS_FB=6;OCS_MINBLOCKS_FWD=6 OCS_FB_RING=2" "cfg3fb" >> gpurun_out/r2e_sum.txt 2>&1
timeout 900 python bench.py --no-configs > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; echo "bench n1 rc=$?" >> gpurun_out/r2e_sum.txt
OCS_B200_STAGER_THREADS=12 timeout 900 python bench.py --no-configs --no-cpu > gpurun_out/r2e_bench_n1_t12.json 2> gpurun_out/r2e_bench_n1_t12.err
cat gpurun_out/r2e_sum.txt
