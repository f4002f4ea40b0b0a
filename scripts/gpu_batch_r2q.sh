#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
for f in 0 1; do
  echo "=== LFM_UNET_FUSE_OUT=$f" >> $O/r2q_fuse.log
  LFM_UNET_FUSE_OUT=$f timeout 300 python scripts/unet_profile.py 32 5 >> $O/r2q_fuse.log 2>&1
  LFM_UNET_FUSE_OUT=$f timeout 300 python scripts/edm_profile.py >> $O/r2q_fuse.log 2>&1
  LFM_UNET_FUSE_OUT=$f timeout 300 python scripts/vae_profile.py 16 >> $O/r2q_fuse.log 2>&1
done
LFM_UNET_FUSE_OUT=1 timeout 900 python -m pytest tests -m gpu -q --timeout=900 -k "unet or edm or vae" > $O/r2q_pytest_fuse.log 2>&1; echo rc=$? >> $O/r2q_pytest_fuse.log
timeout 1500 compute-sanitizer --tool racecheck --print-limit 2000 python scripts/sanitizer_run.py > $O/r2q_racecheck.log 2>&1; echo rc=$? >> $O/r2q_racecheck.log
cat $O/r2q_fuse.log; tail -4 $O/r2q_pytest_fuse.log; tail -5 $O/r2q_racecheck.log
