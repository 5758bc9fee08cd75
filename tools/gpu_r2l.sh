#!/bin/bash
# serpentine tile order A/B: HB_ATT_REVERSE x HB_MLP_REVERSE, 16 regions, 512 patches per launch
set -u
mkdir -p gpurun_out
for cfg in "0 0" "1 0" "0 1" "1 1" "0 0" "1 1"; do
  set -- $cfg
  echo -n "att_reverse=$1 mlp_reverse=$2 "
  HB_ATT_REVERSE=$1 HB_MLP_REVERSE=$2 HB_VIT256_MAX_PATCHES=512 python tools/exp_group_size.py --child 16 2>&1 | tail -1
done | tee gpurun_out/r02_serpentine.txt
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -x -q 2>&1 | tail -2
