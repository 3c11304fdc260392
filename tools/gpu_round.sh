#!/bin/bash
# One GPU-box visit that produces everything kept under profiles/: parity tests, bench (both arms), the other
# configurations, the ncu launch list of the bench command and ncu --set full captures of every kernel of a pass.
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
for P in 1 3 4; do echo "pipes $P: $(EHB_PIPES=$P timeout 300 python bench.py --steps 1000 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'])")"; done | tee $O/pipes.txt
timeout 900 python tools/run_configs.py > $O/configs.jsonl 2> $O/configs.err; cat $O/configs.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu > $O/ncu_launches.log 2>&1
for K in table front raster raster_big windows compose pairgrad; do
  EHB_PIPES=1 EHB_BENCH_NOGRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ehb_k_$K\$ -s 14 -c 1 -f -o $O/$K \
    python bench.py --steps 4 --warmup 3 --no-cpu > $O/ncu_$K.log 2>&1
  ls -la $O/$K.ncu-rep
done
