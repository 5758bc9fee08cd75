#!/bin/bash
# ncu --set full of the CLAM tensor-core score kernel with the gate on the tensor core (1 fold, 5 folds)
set -u
mkdir -p gpurun_out
python tools/bench_clam.py --size hipt_smaller --folds 1 2>&1 | grep '"folds"' | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:clam_scores_tc -s 3 -c 1 -f -o gpurun_out/r02ac_clam_tc_1fold python tools/bench_clam.py --size hipt_smaller --folds 1 > gpurun_out/r02ac_clam1.log 2>&1; echo "clam1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:clam_scores_tc -s 3 -c 1 -f -o gpurun_out/r02ac_clam_tc_5fold python tools/bench_clam.py --size hipt_smaller --folds 5 > gpurun_out/r02ac_clam5.log 2>&1; echo "clam5 rc=$?"
ls -la gpurun_out/r02ac*.ncu-rep
