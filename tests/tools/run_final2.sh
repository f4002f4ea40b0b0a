set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/final2_pytest.log
timeout 120 python scripts/unet_profile.py 32 5 2>&1 | tail -1
LFM_CONV_IN_STRIP=0 timeout 120 python scripts/unet_profile.py 32 5 2>&1 | tail -1
timeout 120 python scripts/edm_profile.py 64 5 2>&1 | tail -1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py 2>/dev/null | tail -1 > gpurun_out/final2_bench.json; cat gpurun_out/final2_bench.json | cut -c1-400
