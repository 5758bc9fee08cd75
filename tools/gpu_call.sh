#!/bin/bash
# usage: tools/gpu_call.sh TAG [tests] [smoke] [bench ARGS...]   (one B200; outputs under gpurun_out/TAG_*)
set -u
mkdir -p gpurun_out
TAG=$1; shift
while [ $# -gt 0 ]; do
  case "$1" in
    tests) python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${TAG}_tests.log; shift;;
    smoke) python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log; shift;;
    bench) shift; python bench.py "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench.err; python tools/show_bench.py gpurun_out/${TAG}_bench.json; break;;
    *) echo "unknown step $1"; exit 2;;
  esac
done
