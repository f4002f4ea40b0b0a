#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r2g_perf.log
for d in 0 1 2 4; do
  echo "=== gemm2, LFM_G2_DBG=$d (0 normal, 1 no drain, 2 TMEM reads only, 4 no global stores)" >> $O/r2g_perf.log
  LFM_PERF_BN=512 LFM_G2_DBG=$d timeout 200 python tests/tools/gpu_bringup.py perf16k 2>&1 | grep "bn=512" >> $O/r2g_perf.log
done
LFM_L2_HINT=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/r2g_bench_l2hint.json 2> $O/r2g_bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/r2g_bench_base.json 2>> $O/r2g_bench.err
cat $O/r2g_perf.log; cut -c1-140 $O/r2g_bench_l2hint.json; cut -c1-140 $O/r2g_bench_base.json
