#!/bin/bash
# where does clam_combine_kernel's time go (hipt_medium: 30 us)?  + the new run-to-run test
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "bit_identical_run_to_run or writes_stay_inside or ragged" 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:clam_combine -s 3 -c 1 -f -o gpurun_out/r02z_combine_medium python tools/bench_clam.py --size hipt_medium --folds 1 > gpurun_out/r02z_combine.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:clam_work_table -s 3 -c 1 -f -o gpurun_out/r02z_table python tools/bench_clam.py --size hipt_medium --folds 1 > gpurun_out/r02z_table.log 2>&1; echo "ncu rc=$?"
