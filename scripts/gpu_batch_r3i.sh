#!/bin/bash
# round 2, session 2: evidence for the mma.sync flash attention kernel - launch list of a DiT-XL/2 evaluation and ncu --set full of the kernel
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 606 -c 202 --csv --log-file $O/r3i_launches_xl2.csv python scripts/dit_profile.py 64 1 "DiT-XL/2" > $O/r3i_ncu_xl2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attention_flash" --launch-skip 10 -c 1 -o $O/r3i_flash_xl2 python scripts/dit_profile.py 64 1 "DiT-XL/2" > $O/r3i_ncu_full_xl2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attention_flash" --launch-skip 10 -c 1 -o $O/r3i_flash_l2_r64 python scripts/dit_profile.py 16 1 "DiT-L/2" 64 > $O/r3i_ncu_full_r64.log 2>&1
ls -la $O/r3i_*
