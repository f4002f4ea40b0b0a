#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -k "unet or edm or vae or cli or smoke or adm" > $O/r2n_pytest.log 2>&1; echo rc=$? >> $O/r2n_pytest.log
timeout 300 python scripts/unet_profile.py 32 5 >> $O/r2n_unet.log 2>&1
timeout 300 python scripts/edm_profile.py >> $O/r2n_unet.log 2>&1
timeout 300 python scripts/vae_profile.py 16 >> $O/r2n_unet.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2n_vae_launches.csv python scripts/vae_profile.py 16 > $O/r2n_vae_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gn_apply_kernel<1>' -c 2 -o $O/r2n_gn_apply1 python scripts/vae_profile.py 16 > $O/r2n_ncu.log 2>&1
ncu -i $O/r2n_gn_apply1.ncu-rep --page raw --csv > $O/r2n_gn_apply1_raw.csv 2>/dev/null
grep -B5 -A25 "Error\|assert" $O/r2n_pytest.log | head -80; tail -6 $O/r2n_pytest.log; cat $O/r2n_unet.log; tail -3 $O/r2n_ncu.log
