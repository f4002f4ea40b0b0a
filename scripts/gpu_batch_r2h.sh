#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r2h_perf.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "test_gemm_epilogues or fixture" --timeout=300 > $O/r2h_test.log 2>&1; echo rc=$? >> $O/r2h_test.log
for d in 0 8; do
  echo "=== gemm2, LFM_G2_DBG=$d (0 = staging + TMA store, packed f32x2 GELU/bias; 8 = bf16 rows stored straight from registers)" >> $O/r2h_perf.log
  LFM_PERF_BN=512 LFM_G2_DBG=$d timeout 200 python tests/tools/gpu_bringup.py perf16k 2>&1 | grep "bn=512" >> $O/r2h_perf.log
done
LFM_G2_DBG=8 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "test_gemm_epilogues" --timeout=300 > $O/r2h_test_direct.log 2>&1; echo rc=$? >> $O/r2h_test_direct.log
LFM_G2_FLAGS=8 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/r2h_bench_direct.json 2> $O/r2h_bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/r2h_bench_base.json 2>> $O/r2h_bench.err
tail -3 $O/r2h_test.log; tail -3 $O/r2h_test_direct.log; cat $O/r2h_perf.log; cut -c1-140 $O/r2h_bench_direct.json; cut -c1-140 $O/r2h_bench_base.json
