#!/bin/bash
# N GPUs of one box: the sharded slide-set bench (strong scaling over the fixed 2,000-region set)
set -u
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err; echo "bench n$N rc=$?"; tail -3 gpurun_out/n${N}_bench.err
python tools/show_bench.py gpurun_out/n${N}_bench.json | cut -c1-600
