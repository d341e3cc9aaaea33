#!/bin/bash
# usage: ab3.sh "<bt list>"   cfg2 at several batch sizes: per-kernel trajectory-steps/s
for bt in $1; do
  python bench.py --config cfg2 --bt $bt --steps 10 --warmup 3 --no-cpu --no-e2e 2>/tmp/ab_err.txt | python -c "
import sys,json
j=json.loads(sys.stdin.read()); ks=j['roofline']['kernels']; bt=$bt
print('bt',bt,'value %.4e'%j['value'],' '.join('%s=%.3fms %.4e/s'%(k['kernel'][4:7],k['ms'],bt*1000/k['ms']*1e3) for k in ks))
"
done
