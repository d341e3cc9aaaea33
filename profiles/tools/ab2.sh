#!/bin/bash
# usage: ab2.sh "<defs1> <defs2> ..." "<configs>"   (defs: OCS_B200_NVRTC_DEFS strings, 'none' for defaults)
for cfg in $2; do
for d in $1; do
  if [ $d = none ]; then unset OCS_B200_NVRTC_DEFS; else export OCS_B200_NVRTC_DEFS="$d"; fi
  python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu --no-e2e 2>/tmp/ab_err.txt | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    j=json.loads(l); ks=j['roofline']['kernels']; print('$cfg','$d','%.3e'%j['value'],' '.join('%s=%.3f'%(k['kernel'][4:22],k['ms']) for k in ks))
except Exception as e: print('$cfg','$d','FAILED',l[:200]); print(open('/tmp/ab_err.txt').read()[-600:])
"
done; done
