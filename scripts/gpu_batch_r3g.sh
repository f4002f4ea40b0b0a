#!/bin/bash
# round 2, session 2, GPU call 6: DhariwalUNet self-attention at 32 x 32 (1024 tokens, flash kernel) against the reference fixture
# edm_attn32, the EDM tests, and the full bench line with the two other-geometry entries (DiT-XL/2, DiT-L/2 on 64x64 latents).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout=300 -k "edm" > $O/r3g_pytest_edm.log 2>&1; echo rc=$? >> $O/r3g_pytest_edm.log; tail -n 4 $O/r3g_pytest_edm.log
timeout 900 python bench.py 2> $O/r3g_bench_full.err | tail -n 1 > $O/r3g_bench_full.json; cut -c1-200 $O/r3g_bench_full.json; tail -n 3 $O/r3g_bench_full.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3g_bench_full.json").read())
for k in ("dit_xl2", "dit_l2_64x64_latents", "cfg3", "cfg4"):
    print(k, d["other_configs"].get(k))
PY
