set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "unet or edm" 2>&1 | tail -2
timeout 120 python scripts/unet_profile.py 32 5 2>&1 | tail -1
LFM_UNET_ATTN_MMA=0 timeout 120 python scripts/unet_profile.py 32 5 2>&1 | tail -1
timeout 120 python scripts/edm_profile.py 64 5 2>&1 | tail -1
LFM_UNET_ATTN_MMA=0 timeout 120 python scripts/edm_profile.py 64 5 2>&1 | tail -1
timeout 120 python scripts/unet_profile.py 8 5 2>&1 | tail -1
