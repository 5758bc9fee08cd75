#!/bin/bash
# A/B of the CLAM score kernel: L2 prefetch ahead of the shared-memory ring (3 / 6 tiles / off), fold split vs tile alternation
set -u
for lib in libhipt_b200 exp_nopf exp_pf6 exp_alt; do
for cfg in "hipt_smaller 1" "hipt_smaller 5" "hipt_small 2" "hipt_medium 1"; do set -- $cfg; echo -n "$lib "; HB_LIB_PATH=$PWD/hipt_abmil_atec23_b200/lib/$lib.so timeout 120 python tools/bench_clam.py --size $1 --folds $2 2>&1 | grep '"folds"' | cut -c1-140; done
done
HB_LIB_PATH=$PWD/hipt_abmil_atec23_b200/lib/libhipt_b200.so timeout 120 python tools/check_clam_tc_accuracy.py hipt_smaller 5 2>&1 | tail -1
