#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:ln_modulate --launch-skip 60 -c 1 -o $O/r2x_ln python scripts/dit_profile.py 64 1 > $O/r2x_ncu.log 2>&1
ncu -i $O/r2x_ln.ncu-rep --page raw --csv > $O/r2x_ln_raw.csv 2>/dev/null
ncu -i $O/r2x_ln.ncu-rep --page source --csv > $O/r2x_ln_source.csv 2>/dev/null
tail -3 $O/r2x_ncu.log; wc -c $O/r2x_ln_raw.csv
