#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -k "unet or edm or vae or cli or smoke or adm" > $O/r2p_pytest.log 2>&1; echo rc=$? >> $O/r2p_pytest.log
timeout 300 python scripts/unet_profile.py 32 5 >> $O/r2p_unet.log 2>&1
timeout 300 python scripts/edm_profile.py >> $O/r2p_unet.log 2>&1
timeout 300 python scripts/vae_profile.py 16 >> $O/r2p_unet.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2p_unet_launches.csv python scripts/unet_profile.py 32 1 > $O/r2p_unet_ncu.log 2>&1
grep -B5 -A25 "Error\|assert" $O/r2p_pytest.log | head -80; tail -6 $O/r2p_pytest.log; cat $O/r2p_unet.log
