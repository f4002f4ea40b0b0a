#!/bin/bash
# round 2, session 2, GPU call 5: the mma.sync attention kernels for head_dim 72 / 1024 tokens (kernel parity, reference fixtures
# mini_xl2 / mini_xl4 / mini_r64p2), then - only if those pass - the whole GPU suite on this build, and DiT-XL/2 + 64x64-latent timings.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_parity.py -q -x --timeout=200 -k "attention_mma_kernels or ((forward_vs_reference_fixture or fixed_step_samplers) and (xl or r64 or p4 or p8 or r16))" > $O/r3f_pytest_new.log 2>&1; rc=$?; echo rc=$rc >> $O/r3f_pytest_new.log; tail -n 12 $O/r3f_pytest_new.log
if [ $rc -ne 0 ]; then echo "new-kernel tests failed: stopping"; exit 0; fi
timeout 300 python scripts/dit_profile.py 64 10 "DiT-XL/2" 2>&1 | tail -n 1 | tee $O/r3f_xl2.log
timeout 300 python scripts/dit_profile.py 16 10 "DiT-L/2" 64 2>&1 | tail -n 1 | tee $O/r3f_l2_r64.log
timeout 1200 python -m pytest tests -x -q -m gpu > $O/r3f_pytest.log 2>&1; echo rc=$? >> $O/r3f_pytest.log; tail -n 3 $O/r3f_pytest.log
