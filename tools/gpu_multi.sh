#!/bin/bash
# N-GPU visit (gpurun --gpus N): peer-mailbox all-reduce vs NCCL, sharded solve, weak-scaling bench lines.
N=${1:-2}
O=gpurun_out/multi$N
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/test_multi_gpu.py > $O/test_multi_gpu.log 2>&1; echo "test_multi_gpu rc=$?"; tail -6 $O/test_multi_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2000 --warmup 20 > $O/bench_peer.json 2> $O/bench_peer.err; echo "bench peer rc=$?"; tail -1 $O/bench_peer.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 2000 --warmup 20 --collective nccl > $O/bench_nccl.json 2> $O/bench_nccl.err; echo "bench nccl rc=$?"; tail -1 $O/bench_nccl.json
timeout 300 python bench.py --steps 2000 --no-cpu > $O/bench_1.json 2>/dev/null; tail -1 $O/bench_1.json
