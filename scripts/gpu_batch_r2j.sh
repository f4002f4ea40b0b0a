#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout=900 > $O/r2j_pytest.log 2>&1; echo rc=$? >> $O/r2j_pytest.log
LFM_LN_FUSE=1 timeout 300 python __graft_entry__.py smoke > $O/r2j_smoke_lnfuse.log 2>&1; rc=$?; echo "rc=$rc" >> $O/r2j_smoke_lnfuse.log
if [ $rc -eq 0 ]; then
  LFM_LN_FUSE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "fixture or properties or odd_batch or cfg_identity" --timeout=600 > $O/r2j_pytest_lnfuse.log 2>&1; echo rc=$? >> $O/r2j_pytest_lnfuse.log
  LFM_LN_FUSE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/r2j_bench_lnfuse.json 2> $O/r2j_bench.err
fi
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/r2j_bench_base.json 2>> $O/r2j_bench.err
for h in 1 0; do
  echo "=== LFM_UNET_H1_BF16=$h" >> $O/r2j_unet.log
  LFM_UNET_H1_BF16=$h timeout 300 python scripts/unet_profile.py 32 5 >> $O/r2j_unet.log 2>&1
  LFM_UNET_H1_BF16=$h timeout 300 python scripts/edm_profile.py >> $O/r2j_unet.log 2>&1
  LFM_UNET_H1_BF16=$h timeout 300 python scripts/vae_profile.py 16 >> $O/r2j_unet.log 2>&1
done
tail -6 $O/r2j_pytest.log; tail -3 $O/r2j_smoke_lnfuse.log; tail -4 $O/r2j_pytest_lnfuse.log; cut -c1-140 $O/r2j_bench_lnfuse.json; cut -c1-140 $O/r2j_bench_base.json; cat $O/r2j_unet.log
