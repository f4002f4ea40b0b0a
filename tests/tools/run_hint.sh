set -x
for h in 0 1 0 1; do LFM_L2_HINT=$h timeout 120 python scripts/dit_profile.py 64 40 2>&1 | tail -1; done
LFM_L2_HINT=1 timeout 100 python -m pytest tests/test_gpu_parity.py -x -q -k "fixture and not unet and not edm" 2>&1 | tail -2
timeout 120 python scripts/unet_profile.py 32 5 2>&1 | tail -1
LFM_PDL=0 timeout 120 python scripts/unet_profile.py 32 5 2>&1 | tail -1
timeout 120 python scripts/edm_profile.py 64 5 2>&1 | tail -1
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
