#!/bin/bash
# full GPU test suite, smoke, default bench line (TAG = $1; "ref" as $2 adds the reference arm)
set -u
mkdir -p gpurun_out
TAG=${1:-r02u}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/${TAG}_bench.json
if [ "${2:-}" = "ref" ]; then
timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench_reference.json
fi
