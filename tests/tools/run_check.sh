set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "unet or edm" 2>&1 | tail -2
timeout 120 python scripts/unet_profile.py 32 5 2>&1 | tail -1
timeout 120 python scripts/edm_profile.py 64 5 2>&1 | tail -1
timeout 120 python scripts/unet_profile.py 8 5 2>&1 | tail -1
timeout 600 python bench.py 2>/dev/null | tail -1 > gpurun_out/check_bench.json; python -c "
import json; d=json.load(open('gpurun_out/check_bench.json')); print(d['value'], d['e2e']['value'], d['roofline'], d['clocks'])"
