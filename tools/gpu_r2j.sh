#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ingest.py -m gpu -x -q -s > gpurun_out/r2j_ingest_tests.log 2>&1; echo "ingest tests rc=$?"; grep -E "backend|cosine|passed|failed|Error" gpurun_out/r2j_ingest_tests.log | tail -12
python tools/bench_ingest.py > gpurun_out/r2j_ingest.json 2> gpurun_out/r2j_ingest.err; echo "bench_ingest rc=$?"; tail -3 gpurun_out/r2j_ingest.err; cat gpurun_out/r2j_ingest.json
python tools/bench_ingest.py --subsampling 2 > gpurun_out/r2j_ingest_420.json 2>> gpurun_out/r2j_ingest.err; cat gpurun_out/r2j_ingest_420.json
