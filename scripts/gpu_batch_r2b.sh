#!/bin/bash
# one gpurun call: smoke of the LayerNorm-finisher path first (falls back to LFM_LN_FUSE=0 for the rest if it fails)
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python __graft_entry__.py smoke > $O/r2b_smoke.log 2>&1; rc=$?
echo "smoke rc=$rc" >> $O/r2b_smoke.log
if [ $rc -ne 0 ]; then export LFM_LN_FUSE=0; echo "LN fusion disabled for the rest" >> $O/r2b_smoke.log; fi
timeout 2000 python -m pytest tests -m gpu -q --timeout=900 > $O/r2b_pytest.log 2>&1; echo rc=$? >> $O/r2b_pytest.log
timeout 300 python tests/tools/gpu_bringup.py attn 2>&1 | grep "variant [35]:" > $O/r2b_attn.log
timeout 600 python scripts/vae_profile.py 1 16 64 > $O/r2b_vae.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2b_bench.json 2> $O/r2b_bench.err
LFM_LN_FUSE=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/r2b_bench_nofuse.json 2>> $O/r2b_bench.err
timeout 900 ncu --set full --clock-control none -k regex:"nvjet|cutlass|gemm2|cublas|sm100" -c 24 -o $O/r2b_gemm python scripts/gemm_profile.py 3 > $O/r2b_gemm_ncu.log 2>&1
tail -3 $O/r2b_smoke.log; tail -15 $O/r2b_pytest.log; cat $O/r2b_attn.log $O/r2b_vae.log; cut -c1-300 $O/r2b_bench.json; cut -c1-300 $O/r2b_bench_nofuse.json; tail -3 $O/r2b_gemm_ncu.log
