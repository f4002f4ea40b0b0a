#!/bin/bash
# round 2, session 2, GPU call 2: packed-f32x2 arithmetic in the LayerNorm pass (LFM_LN_X2) and the attention softmax (LFM_ATTN_X2),
# and the persistent 32-warp LayerNorm without the shared-memory ring (LFM_LN_ROWS=32): same-box bench A/B, per-launch times, parity.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/r3c_ab.log
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras 2>> $O/r3c_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])" >> $O/r3c_ab.log
}
run base LFM_LN_X2=0
run ln_x2 LFM_LN_X2=1
run ln32_x2 LFM_LN_ROWS=32 LFM_LN_X2=1
run attn_x2 LFM_ATTN_X2=1
run base LFM_LN_X2=0
run ln_x2+attn_x2 LFM_LN_X2=1 LFM_ATTN_X2=1
cat $O/r3c_ab.log
for cfg in "LFM_LN_X2=0" "LFM_LN_X2=1 LFM_ATTN_X2=1" "LFM_LN_ROWS=32 LFM_LN_X2=1" "LFM_LN_ROWS=32 LFM_LN_X2=0"; do
  tag=$(echo $cfg | tr ' =' '__')
  env $cfg timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ln_modulate|attention3" --launch-skip 60 -c 9 --csv --log-file $O/r3c_k_$tag.csv python scripts/dit_profile.py 64 1 > /dev/null 2>&1
  echo "== $cfg" >> $O/r3c_kernels.log
  grep duration $O/r3c_k_$tag.csv | awk -F'","' '{print $5, $NF}' | sed 's/(.*)//' >> $O/r3c_kernels.log
done
cat $O/r3c_kernels.log
LFM_LN_X2=1 LFM_ATTN_X2=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout=600 \
  -k "(test_attention or forward_vs_reference_fixture or fixed_step_samplers or full_size_dit or odd_batch or cfg_identity or full_size_properties_dit or edm_forward) and not unet and not vae" \
  > $O/r3c_pytest_x2.log 2>&1; echo rc=$? >> $O/r3c_pytest_x2.log; tail -3 $O/r3c_pytest_x2.log
LFM_LN_ROWS=32 LFM_LN_X2=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout=600 \
  -k "(forward_vs_reference_fixture or fixed_step_samplers or full_size_dit or odd_batch) and not unet and not edm and not vae" \
  > $O/r3c_pytest_ln32.log 2>&1; echo rc=$? >> $O/r3c_pytest_ln32.log; tail -3 $O/r3c_pytest_ln32.log
