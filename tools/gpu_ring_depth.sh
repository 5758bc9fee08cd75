#!/bin/bash
# ring-depth sensitivity of the CLAM score kernel: one hipt_smaller fold (12 stages fit) with the ring capped at 4 / 5 / 6 / 8 stages
set -u
for lib in exp_st4 exp_st5 exp_st6 exp_st8 libhipt_b200; do echo -n "$lib "; HB_LIB_PATH=$PWD/hipt_abmil_atec23_b200/lib/$lib.so timeout 120 python tools/bench_clam.py --size hipt_smaller --folds 1 2>&1 | grep '"folds"' | cut -c1-150; done
