#!/bin/bash
# round 2, session 2, GPU call 3: attention with S issued as two key halves, the first one early (LFM_ATTN_SPLIT): kernel parity first
# (short timeout: new barrier protocol), per-launch times, same-box bench A/B; one run of the direct-store bf16 epilogue (LFM_G2_FLAGS=8).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/r3d_ab.log $O/r3d_kernels.log
LFM_ATTN_SPLIT=1 timeout 150 python -m pytest tests/test_gpu_parity.py -q -x --timeout=120 -k "test_attention" > $O/r3d_pytest_attn.log 2>&1; rc=$?; echo rc=$rc >> $O/r3d_pytest_attn.log; tail -n 3 $O/r3d_pytest_attn.log
if [ $rc -ne 0 ]; then echo "attention split variant failed its kernel test: stopping"; exit 0; fi
LFM_ATTN_SPLIT=1 LFM_ATTN_X2=1 timeout 400 python -m pytest tests/test_gpu_parity.py -q -x --timeout=300 \
  -k "(test_attention or forward_vs_reference_fixture or fixed_step_samplers or full_size_dit or odd_batch or cfg_identity or full_size_properties_dit or edm_forward) and not unet and not vae" \
  > $O/r3d_pytest.log 2>&1; echo rc=$? >> $O/r3d_pytest.log; tail -n 3 $O/r3d_pytest.log
for cfg in "LFM_ATTN_SPLIT=0" "LFM_ATTN_SPLIT=1" "LFM_ATTN_SPLIT=1 LFM_ATTN_X2=1"; do
  tag=$(echo $cfg | tr ' =' '__')
  env $cfg timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"attention3" --launch-skip 20 -c 6 --csv --log-file $O/r3d_k_$tag.csv python scripts/dit_profile.py 64 1 > /dev/null 2>&1
  echo "== $cfg" >> $O/r3d_kernels.log
  grep duration $O/r3d_k_$tag.csv | awk -F'","' '{print $5, $NF}' | sed 's/(.*)//' >> $O/r3d_kernels.log
done
cat $O/r3d_kernels.log
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras 2>> $O/r3d_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])" >> $O/r3d_ab.log
}
run base LFM_ATTN_SPLIT=0
run split LFM_ATTN_SPLIT=1
run split+x2 LFM_ATTN_SPLIT=1 LFM_ATTN_X2=1 LFM_LN_X2=1
run base LFM_ATTN_SPLIT=0
run split+x2 LFM_ATTN_SPLIT=1 LFM_ATTN_X2=1 LFM_LN_X2=1
run g2_direct_bf16 LFM_G2_FLAGS=8
cat $O/r3d_ab.log
