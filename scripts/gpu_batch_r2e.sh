#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r2e_perf.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "test_gemm_epilogues and 640" --timeout=120 > $O/r2e_gemm_test.log 2>&1; echo rc=$? >> $O/r2e_gemm_test.log
for sk in 0 1 2; do
  echo "=== coalesced-store epilogue, skew $sk" >> $O/r2e_perf.log
  LFM_G5_SKEW=$sk timeout 200 python tests/tools/gpu_bringup.py perf16k 2>&1 | grep "bn=640" >> $O/r2e_perf.log
done
echo "=== TMA-store epilogue (dbg 2), skew 2" >> $O/r2e_perf.log
LFM_G5_DBG=2 timeout 200 python tests/tools/gpu_bringup.py perf16k 2>&1 | grep "bn=640" >> $O/r2e_perf.log
LFM_BN_QKV=640 LFM_BN_FC1=640 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/r2e_bench_640qkvfc1.json 2> $O/r2e_bench.err
LFM_BN_QKV=640 LFM_BN_PROJ=640 LFM_BN_FC1=640 LFM_BN_FC2=640 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/r2e_bench_640all.json 2>> $O/r2e_bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/r2e_bench_512.json 2>> $O/r2e_bench.err
tail -3 $O/r2e_gemm_test.log; cat $O/r2e_perf.log; for f in $O/r2e_bench_640qkvfc1.json $O/r2e_bench_640all.json $O/r2e_bench_512.json; do cut -c1-140 $f; done
