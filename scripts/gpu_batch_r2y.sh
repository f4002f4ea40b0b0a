#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
for r in 16 0 16 0; do
  LFM_LN_ROWS=$r timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras 2>> $O/r2y_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('LN_ROWS=$r', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])" >> $O/r2y_ab.log
done
LFM_LN_ROWS=0 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "fixture or properties or odd_batch or cfg_identity" --timeout=600 > $O/r2y_pytest0.log 2>&1; echo rc=$? >> $O/r2y_pytest0.log
LFM_LN_ROWS=0 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ln_modulate --launch-skip 60 -c 4 --csv --log-file $O/r2y_ln0.csv python scripts/dit_profile.py 64 1 > /dev/null 2>&1
cat $O/r2y_ab.log; tail -3 $O/r2y_pytest0.log; grep duration $O/r2y_ln0.csv | cut -d, -f5,15- | head -5
