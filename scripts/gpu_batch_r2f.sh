#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/r2f_perf.log
echo "=== gemm2 (256x256 pair tile), normal" >> $O/r2f_perf.log
LFM_PERF_BN=512 timeout 200 python tests/tools/gpu_bringup.py perf16k 2>&1 | grep "bn=512" >> $O/r2f_perf.log
echo "=== gemm2, epilogue does not drain TMEM (LFM_G2_DBG=1)" >> $O/r2f_perf.log
LFM_PERF_BN=512 LFM_G2_DBG=1 timeout 200 python tests/tools/gpu_bringup.py perf16k 2>&1 | grep "bn=512" >> $O/r2f_perf.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitizer_run.py > $O/r2f_sanitizer_memcheck.log 2>&1; echo "exit code $?" >> $O/r2f_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitizer_run.py > $O/r2f_sanitizer_racecheck.log 2>&1; echo "exit code $?" >> $O/r2f_sanitizer_racecheck.log
timeout 2400 python -m pytest tests -m gpu -q --timeout=900 > $O/r2f_pytest.log 2>&1; echo rc=$? >> $O/r2f_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2f_launches_dit.csv python scripts/dit_profile.py 64 3 > $O/r2f_ncu_dit.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r2f_launches_unet.csv python scripts/unet_profile.py 32 1 > $O/r2f_ncu_unet.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r2f_launches_edm.csv python scripts/edm_profile.py 64 1 > $O/r2f_ncu_edm.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2f_launches_vae.csv python scripts/vae_profile.py 16 > $O/r2f_ncu_vae.log 2>&1
cat $O/r2f_perf.log; tail -4 $O/r2f_sanitizer_memcheck.log; tail -4 $O/r2f_sanitizer_racecheck.log; tail -6 $O/r2f_pytest.log
