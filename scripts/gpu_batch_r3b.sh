#!/bin/bash
# round 2, session 2, GPU call 2 (final evidence): same-box A/B/A/B of the packed-f32x2 LayerNorm (LFM_LN_X2), then - with the
# faster setting exported - the whole GPU suite, smoke(), the full bench line and the ncu launch list of one bench step.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/r3b_ab.log
for v in 0 1 0 1; do
  LFM_LN_X2=$v timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras 2>> $O/r3b_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('LN_X2=$v', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])" >> $O/r3b_ab.log
done
cat $O/r3b_ab.log
X2=$(python - <<'PY'
import re
v = {0: [], 1: []}
for line in open("gpurun_out/r3b_ab.log"):
    m = re.match(r"LN_X2=(\d) ([\d.]+)", line)
    if m:
        v[int(m.group(1))].append(float(m.group(2)))
ok = all(len(x) == 2 for x in v.values())
print(1 if ok and min(v[1]) > max(v[0]) else 0)
PY
)
echo "chosen LFM_LN_X2=$X2" | tee $O/r3b_choice.log
export LFM_LN_X2=$X2
timeout 1200 python -m pytest tests -x -q -m gpu > $O/r3b_pytest.log 2>&1; echo rc=$? >> $O/r3b_pytest.log; tail -3 $O/r3b_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/r3b_smoke.log
timeout 900 python bench.py 2> $O/r3b_bench_full.err | tail -1 > $O/r3b_bench_full.json; cut -c1-400 $O/r3b_bench_full.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 600 -c 400 --csv --log-file $O/r3b_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/r3b_ncu_bench.log 2>&1
echo finished
