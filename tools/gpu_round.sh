#!/bin/bash
# Round evidence on one B200: GPU parity tests, smoke, the default bench line, the reference arm, then ncu (launch
# list + one full capture of the three dominant kernels) on a short bench of the same workload.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/${TAG}_bench.json
if [ "${2:-}" = "ncu" ]; then
SHORT="python bench.py --steps 1 --warmup 3 --regions-per-step 2 --no-cpu-baseline"
$SHORT > gpurun_out/${TAG}_short.json 2> gpurun_out/${TAG}_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 232 -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv $SHORT > gpurun_out/${TAG}_ncu_ll.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'mlp_fused_kernel|attention_tc_kernel' -s 8 -c 2 -o gpurun_out/${TAG}_full $SHORT > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16_kernel' -s 14 -c 2 -o gpurun_out/${TAG}_full_gemm $SHORT > gpurun_out/${TAG}_ncu_full_gemm.log 2>&1
echo "ncu full gemm rc=$?"
fi
