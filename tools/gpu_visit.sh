#!/bin/bash
# One GPU-box visit: parity tests, bench lines (both workloads, both arms), optional extras.
# usage (on the box, from the repo root): bash tools/gpu_visit.sh TAG [sanitize] [configs] [ncu]
TAG=${1:-visit}; shift
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; tail -3 $O/bench.err
timeout 600 python bench.py --workload inview --no-proxy > $O/bench_inview.json 2> $O/bench_inview.err; echo "inview rc=$?"; cat $O/bench_inview.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; cat $O/bench_reference.json
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_k20.json 2>> $O/bench.err; cat $O/bench_k20.json
for A in "$@"; do
  case $A in
  sanitize)
    timeout 900 compute-sanitizer --tool memcheck --log-file $O/memcheck_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > $O/memcheck_smoke.out 2>&1; echo "memcheck smoke rc=$?"; tail -3 $O/memcheck_smoke.log
    timeout 900 compute-sanitizer --tool racecheck --log-file $O/racecheck_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > $O/racecheck_smoke.out 2>&1; echo "racecheck smoke rc=$?"; tail -3 $O/racecheck_smoke.log
    timeout 1200 compute-sanitizer --tool memcheck --log-file $O/memcheck_fused.log python -m pytest tests/test_gpu_parity.py -x -q -k "fused_views_match_oracle and 120" > $O/memcheck_fused.out 2>&1; echo "memcheck fused rc=$?"; tail -3 $O/memcheck_fused.log
    ;;
  configs)
    timeout 1200 python tools/run_configs.py > $O/configs.jsonl 2> $O/configs.err; cat $O/configs.jsonl | cut -c1-600
    ;;
  ncu)
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
        python bench.py --steps 4 --warmup 3 --no-cpu > $O/ncu_launches.log 2>&1
    for K in ${NCU_KERNELS:-front raster raster_big tiles}; do
      EHB_PIPES=1 EHB_BENCH_NOGRAPH=1 EHB_VALUE_SLOTS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ehb_k_$K\$ -s 10 -c 1 -f -o $O/$K \
        python bench.py --steps 4 --warmup 3 --no-cpu > $O/ncu_$K.log 2>&1
      ls -la $O/$K.ncu-rep
    done
    ;;
  esac
done
ls -la $O
