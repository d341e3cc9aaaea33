#!/bin/bash
# usage: ab.sh "<variants>" "<configs>"  -> one line per (variant, config): kernel ms
V=optimal-control-solvers_b200/build/variants   # build.py --variant output
for cfg in $2; do
for v in $1; do
  if [ $v = default ]; then unset OCS_B200_LIB; else export OCS_B200_LIB=$PWD/$V/libocs_$v.so; fi
  python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu --no-e2e 2>/tmp/ab_err.txt | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    j=json.loads(l); ks=j['roofline']['kernels']; print('$cfg','$v','%.3e'%j['value'],' '.join('%s=%.3f'%(k['kernel'][4:22],k['ms']) for k in ks))
except Exception as e: print('$cfg','$v','FAILED',l[:200])
"
done; done
