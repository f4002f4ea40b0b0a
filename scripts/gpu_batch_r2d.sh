#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
for sk in 0 1 2; do
  echo "=== skew $sk" >> $O/r2d_perf.log
  LFM_G5_SKEW=$sk timeout 200 python tests/tools/gpu_bringup.py perf16k 2>&1 | grep "bn=640" >> $O/r2d_perf.log
done
echo "=== skew 1, no epilogue drain" >> $O/r2d_perf.log
LFM_G5_SKEW=1 LFM_G5_DBG=1 timeout 200 python tests/tools/gpu_bringup.py perf16k 2>&1 | grep "bn=640" >> $O/r2d_perf.log
echo "=== skew 0, no epilogue drain" >> $O/r2d_perf.log
LFM_G5_SKEW=0 LFM_G5_DBG=1 timeout 200 python tests/tools/gpu_bringup.py perf16k 2>&1 | grep "bn=640" >> $O/r2d_perf.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "vae or cli" --timeout=600 > $O/r2d_pytest.log 2>&1; echo rc=$? >> $O/r2d_pytest.log
timeout 300 python __graft_entry__.py smoke > $O/r2d_smoke.log 2>&1; echo rc=$? >> $O/r2d_smoke.log
LFM_UNET_GEMM5=1 timeout 300 python scripts/unet_profile.py 32 5 > $O/r2d_unet_g5.log 2>&1
LFM_UNET_GEMM5=1 timeout 300 python scripts/vae_profile.py 16 >> $O/r2d_unet_g5.log 2>&1
cat $O/r2d_perf.log; tail -5 $O/r2d_pytest.log; tail -3 $O/r2d_smoke.log; cat $O/r2d_unet_g5.log
