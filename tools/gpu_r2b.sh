#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2b_tests.log
python tools/run_kernel.py attention 20 512 2>&1 | tail -2
HB_ATTENTION_V1=1 python tools/run_kernel.py attention 20 512 2>&1 | tail -2
HB_ATTENTION_V1=1 python bench.py --regions 100 --steps 2 --ref-regions 1 > gpurun_out/r2b_bench_v1.json 2> gpurun_out/r2b_bench_v1.err; echo "bench v1 rc=$?"; tail -5 gpurun_out/r2b_bench_v1.err
python tools/show_bench.py gpurun_out/r2b_bench_v1.json
python bench.py --regions 100 --steps 2 --no-cpu-baseline --no-sections > gpurun_out/r2b_bench_v2.json 2> gpurun_out/r2b_bench_v2.err; echo "bench v2 rc=$?"; tail -5 gpurun_out/r2b_bench_v2.err
python tools/show_bench.py gpurun_out/r2b_bench_v2.json
