#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2o_unet_launches.csv python scripts/unet_profile.py 32 1 > $O/r2o_unet_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2o_edm_launches.csv python scripts/edm_profile.py > $O/r2o_edm_ncu.log 2>&1
tail -2 $O/r2o_unet_ncu.log $O/r2o_edm_ncu.log
