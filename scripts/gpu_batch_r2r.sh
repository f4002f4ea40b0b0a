#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum -k regex:ln_modulate --launch-skip 150 -c 8 --csv --log-file $O/r2r_ln_l2.csv python scripts/dit_profile.py 64 3 > $O/r2r_ln.log 2>&1
tail -3 $O/r2r_ln.log
