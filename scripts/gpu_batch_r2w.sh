#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
for r in 8 16 8 16; do
  LFM_LN_ROWS=$r timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras 2>> $O/r2w_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('LN_ROWS=$r', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])" >> $O/r2w_ab.log
done
LFM_LN_ROWS=16 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "fixture or properties or odd_batch or cfg_identity" --timeout=600 > $O/r2w_pytest16.log 2>&1; echo rc=$? >> $O/r2w_pytest16.log
cat $O/r2w_ab.log; tail -3 $O/r2w_pytest16.log
