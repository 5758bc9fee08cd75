#!/bin/bash
# round-2 ncu evidence of the hot path: launch list of a short bench, then --set full of the three top kernels inside the same command
set -u
mkdir -p gpurun_out
CMD="python bench.py --regions 16 --steps 1 --warmup 3 --no-cpu-baseline --no-sections --secondary-steps 0"
$CMD > gpurun_out/ncu2_plain.json 2> gpurun_out/ncu2_plain.err; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02e_launches.csv $CMD > gpurun_out/ncu2_launch.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_fused -s 12 -c 1 -f -o gpurun_out/r02e_mlp_fused $CMD > gpurun_out/ncu2_mlp.log 2>&1; echo "mlp rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc2 -s 12 -c 1 -f -o gpurun_out/r02e_attention $CMD > gpurun_out/ncu2_att.log 2>&1; echo "att rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 25 -c 2 -f -o gpurun_out/r02e_gemm $CMD > gpurun_out/ncu2_gemm.log 2>&1; echo "gemm rc=$?"
ls -la gpurun_out/r02e_*
