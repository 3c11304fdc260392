#!/bin/bash
# quick experiment visit: bench variants (no CPU legs).  usage: gpu_quick.sh TAG [variant-suffix ...]
TAG=${1:-q}; shift
O=gpurun_out/$TAG; mkdir -p $O
summ() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f ms/step %.4f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']), {k: round(v,1) for k,v in d['roofline']['kernel_us'].items()})
    else: print(l.rstrip()[-300:])"; }
echo "== default"; timeout 300 python bench.py --steps 1000 --no-cpu 2>&1 | summ
echo "== default inview"; timeout 300 python bench.py --steps 1000 --no-cpu --workload inview 2>&1 | summ
for V in "$@"; do
  echo "== $V"; EHB_LIB=$PWD/easyhec_b200/libehb_$V.so timeout 300 python bench.py --steps 1000 --no-cpu 2>&1 | summ
done
if [ -f easyhec_b200/libehb_stats.so ]; then
  EHB_LIB=$PWD/easyhec_b200/libehb_stats.so timeout 120 python tools/tile_stats.py headline 2>&1 | tail -3
  EHB_LIB=$PWD/easyhec_b200/libehb_stats.so timeout 120 python tools/tile_stats.py inview 2>&1 | tail -3
fi
