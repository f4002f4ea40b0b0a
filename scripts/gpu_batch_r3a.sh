#!/bin/bash
# round 2, session 2, GPU call 1: new DiT geometries + the tile-starved half-width GEMM mode (parity), small-batch latency A/B,
# launch lists at batch 16 (the per-GPU share of cfg5's strong-scaling run at 8 GPUs) and batch 1, headline sanity.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1000 python -m pytest tests/test_gpu_parity.py -q -x --timeout=600 \
  -k "(gemm_epilogues or test_attention or forward_vs_reference_fixture or fixed_step_samplers or heun_reference or full_size_dit or odd_batch or torchdiffeq_euler or dopri5_with_cfg or cfg_identity or full_size_properties_dit) and not unet and not edm and not vae" \
  > $O/r3a_pytest.log 2>&1; echo rc=$? >> $O/r3a_pytest.log
for h in 1 0 1 0; do
  LFM_GEMM_HALVES=$h timeout 300 python scripts/dit_latency.py "DiT-L/2" 20 5 1,2,4,8,16 2>> $O/r3a_lat.err | sed "s/$/ HALVES=$h/" >> $O/r3a_latency.log
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 531 -c 177 --csv --log-file $O/r3a_launches_b16.csv python scripts/dit_profile.py 16 1 > $O/r3a_ncu_b16.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 531 -c 177 --csv --log-file $O/r3a_launches_b1.csv python scripts/dit_profile.py 1 1 > $O/r3a_ncu_b1.log 2>&1
timeout 300 python scripts/dit_profile.py 64 20 "DiT-L/4" > $O/r3a_dit_l4.log 2>&1
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/r3a_bench.json 2> $O/r3a_bench.err
tail -4 $O/r3a_pytest.log; cat $O/r3a_latency.log; tail -2 $O/r3a_dit_l4.log; cut -c1-200 $O/r3a_bench.json
