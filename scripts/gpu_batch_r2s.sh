#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I lfm_b200/csrc tests/tools/l2_residency.cu -o /tmp/l2res > $O/r2s_build.log 2>&1
timeout 120 /tmp/l2res > $O/r2s_l2res.log 2>&1; echo rc=$? >> $O/r2s_l2res.log
cat $O/r2s_l2res.log
