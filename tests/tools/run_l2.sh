set -x
for z in 0 1; do LFM_ZIGZAG=$z timeout 120 python scripts/dit_profile.py 64 30 2>&1 | tail -1; done
LFM_ZIGZAG=0 LFM_L2_PERSIST_MB=64 LFM_L2_HIT_PCT=100 timeout 120 python scripts/dit_profile.py 64 30 2>&1 | tail -2
LFM_ZIGZAG=1 LFM_L2_PERSIST_MB=64 LFM_L2_HIT_PCT=100 timeout 120 python scripts/dit_profile.py 64 30 2>&1 | tail -2
LFM_ZIGZAG=1 LFM_L2_PERSIST_MB=48 LFM_L2_HIT_PCT=60 timeout 120 python scripts/dit_profile.py 64 30 2>&1 | tail -2
LFM_ZIGZAG=1 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "fixture or full_size or solvers" 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --launch-skip 531 -c 360 --csv --log-file gpurun_out/warm_launches_z0.csv python scripts/dit_profile.py 64 3 > gpurun_out/warm_z0.log 2>&1
LFM_ZIGZAG=1 timeout 300 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --launch-skip 531 -c 360 --csv --log-file gpurun_out/warm_launches_z1.csv python scripts/dit_profile.py 64 3 > gpurun_out/warm_z1.log 2>&1
tail -2 gpurun_out/warm_z1.log
