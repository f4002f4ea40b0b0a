set -x
timeout 120 python scripts/unet_profile.py 32 5 2>&1 | tail -1
LFM_PDL=0 timeout 120 python scripts/unet_profile.py 32 5 2>&1 | tail -1
timeout 120 python scripts/edm_profile.py 64 5 2>&1 | tail -1
LFM_PDL=0 timeout 120 python scripts/edm_profile.py 64 5 2>&1 | tail -1
timeout 120 python scripts/unet_profile.py 8 5 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "unet or edm" 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 584 -c 300 --csv --log-file gpurun_out/edm_launches.csv python scripts/edm_profile.py 64 1 > gpurun_out/edm_ncu.log 2>&1
tail -1 gpurun_out/edm_ncu.log
