#!/bin/bash
# Weak-scaling visit on one box (gpurun --gpus N): the multi-GPU tests, then the bench line at 1, 2, 4 ... N GPUs and the
# exploration workload at 1 and N GPUs.  usage: bash tools/gpu_scale.sh TAG N
TAG=${1:-scale}; N=${2:-8}
O=gpurun_out/$TAG; mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $O/smi.txt 2>&1
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > $O/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 $O/pytest_multi.log
timeout 300 python bench.py --steps 1024 --warmup 20 --no-cpu > $O/bench_1.json 2> $O/bench_1.err; tail -1 $O/bench_1.json | cut -c1-180
G=2
while [ $G -le $N ]; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29520 + G)) bench.py --gpus $G --steps 1024 --warmup 20 > $O/bench_$G.json 2> $O/bench_$G.err; echo "bench $G rc=$?"; tail -1 $O/bench_$G.json | cut -c1-180
  G=$((G * 2))
done
timeout 300 python bench.py --workload explore --steps 5 --no-cpu > $O/explore_1.json 2> $O/explore_1.err; tail -1 $O/explore_1.json | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 bench.py --workload explore --gpus $N --steps 5 > $O/explore_$N.json 2> $O/explore_$N.err; echo "explore $N rc=$?"; tail -1 $O/explore_$N.json | cut -c1-200
python - <<PY
import json, glob, os
v = {}
for f in sorted(glob.glob("$O/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); v[d["n_gpus"]] = d
    except Exception as e: print(f, "unreadable", e)
for n in sorted(v):
    d = v[n]; print("N=%d value %.0f (%.3f of N x 1-GPU) e2e %.0f  %.1f us/step  %s" % (n, d["value"], d["value"] / (n * v[1]["value"]) if 1 in v else 0, d["e2e"]["value"], 1e3 * d["ms_per_step"], d["config"]["collective"][:40]))
PY
