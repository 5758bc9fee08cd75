#!/bin/bash
# two GPUs of one box: two-device tests, then the sharded slide-set bench on 2 ranks
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi_device.py -m gpu -x -q > gpurun_out/n2_tests.log 2>&1; echo "multi-device tests rc=$?"; tail -5 gpurun_out/n2_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --regions ${1:-200} --steps 2 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/n2_bench.err
python tools/show_bench.py gpurun_out/n2_bench.json
