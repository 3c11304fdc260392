#!/bin/bash
# variant comparison: in-flight value / serial / e2e and stage times per library variant.  usage: gpu_var.sh TAG variant...
TAG=${1:-var}; shift
O=gpurun_out/$TAG; mkdir -p $O
summ() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f serial %.0f e2e %.0f' % (d['value'], d['serial']['value'], d['e2e']['value']), {k: round(v,1) for k,v in d['roofline']['kernel_us'].items()})
    else: print(l.rstrip()[-300:])"; }
echo "== default"; timeout 300 python bench.py --steps 1000 --no-cpu 2>&1 | summ
for V in "$@"; do
  echo "== $V"; EHB_LIB=$PWD/easyhec_b200/libehb_$V.so timeout 300 python bench.py --steps 1000 --no-cpu 2>&1 | summ
done
