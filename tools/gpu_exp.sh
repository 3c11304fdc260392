#!/bin/bash
# experiment visit: parity first, then bench variants.  usage: gpu_exp.sh TAG "variant variant ..." [ncu kernels]
TAG=${1:-exp}
O=gpurun_out/$TAG
mkdir -p $O
summ() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f e2e %.0f' % (d['value'], d['e2e']['value']), {k: round(v,1) for k,v in d['roofline']['kernel_us'].items()})
    else: print(l.rstrip())"; }
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
echo "== default"; timeout 300 python bench.py --steps 1000 --no-cpu 2>&1 | summ
for V in $2; do
  echo "== $V"; EHB_LIB=$PWD/easyhec_b200/libehb_$V.so timeout 300 python bench.py --steps 1000 --no-cpu 2>&1 | summ
done
echo "== pipes 2"; EHB_PIPES=2 timeout 300 python bench.py --steps 1000 --no-cpu 2>&1 | summ
for K in $3; do
  EHB_PIPES=1 EHB_BENCH_NOGRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ehb_k_$K\$ -s 14 -c 1 -f -o $O/$K \
    python bench.py --steps 4 --warmup 3 --no-cpu > $O/ncu_$K.log 2>&1
done
ls -la $O
