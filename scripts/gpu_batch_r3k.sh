#!/bin/bash
# round 2, session 2: sanity of the last build after the half-tile rule went back to 2 x tiles <= CTA pairs
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_parity.py -q -x --timeout=150 \
  -k "(gemm_epilogues or forward_vs_reference_fixture or fixed_step_samplers or odd_batch) and not unet and not edm and not vae" \
  > $O/r3k_pytest.log 2>&1; echo rc=$? >> $O/r3k_pytest.log; tail -n 3 $O/r3k_pytest.log
timeout 100 python scripts/dit_latency.py "DiT-L/2" 20 3 1,16 2>/dev/null | tee $O/r3k_latency.log
