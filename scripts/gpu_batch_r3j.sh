#!/bin/bash
# round 2, session 2: launches of at most one wave of tiles entirely in half tiles (the residual GEMMs of the 8-GPU strong-scaling share)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/r3j_latency.log
timeout 400 python -m pytest tests/test_gpu_parity.py -q -x --timeout=300 \
  -k "(gemm_epilogues or forward_vs_reference_fixture or fixed_step_samplers or full_size_dit or odd_batch or cfg_identity) and not unet and not edm and not vae" \
  > $O/r3j_pytest.log 2>&1; echo rc=$? >> $O/r3j_pytest.log; tail -n 3 $O/r3j_pytest.log
for h in 1 0 1 0; do
  LFM_GEMM_HALVES=$h timeout 200 python scripts/dit_latency.py "DiT-L/2" 20 5 8,16,18 2>> $O/r3j_lat.err | sed "s/$/ HALVES=$h/" >> $O/r3j_latency.log
done
cat $O/r3j_latency.log
