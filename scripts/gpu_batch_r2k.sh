#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -k "unet or edm or vae or fid or cli or smoke or adm" > $O/r2k_pytest.log 2>&1; echo rc=$? >> $O/r2k_pytest.log
timeout 300 python scripts/unet_profile.py 32 5 >> $O/r2k_unet.log 2>&1
timeout 300 python scripts/edm_profile.py >> $O/r2k_unet.log 2>&1
timeout 300 python scripts/vae_profile.py 16 >> $O/r2k_unet.log 2>&1
for b in 1 2 4 8 16 32; do timeout 200 python scripts/dit_profile.py $b 20 >> $O/r2k_dit_small.log 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2k_vae_launches.csv python scripts/vae_profile.py 16 > $O/r2k_vae_ncu.log 2>&1
tail -6 $O/r2k_pytest.log; cat $O/r2k_unet.log $O/r2k_dit_small.log
