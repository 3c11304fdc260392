#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list, ncu --set full of the two dominant kernels.
# usage (from the repo root, on the box): bash tools/gpu_round.sh [tag]
TAG=${1:-run}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
cat $O/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
cat $O/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu > $O/ncu_launches.log 2>&1
for K in raster tiles front; do
  EHB_PIPES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ehb_k_$K\$ -s 14 -c 1 -f -o $O/$K \
    python bench.py --steps 4 --warmup 3 --no-cpu > $O/ncu_$K.log 2>&1
  ls -la $O/$K.ncu-rep
done
