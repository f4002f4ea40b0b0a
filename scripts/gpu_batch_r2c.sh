#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "test_gemm_epilogues" --timeout=120 > $O/r2c_gemm_test.log 2>&1; echo rc=$? >> $O/r2c_gemm_test.log
timeout 300 python tests/tools/gpu_bringup.py perf > $O/r2c_perf.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q --timeout=900 --deselect tests/test_gpu_parity.py::test_gemm_epilogues > $O/r2c_pytest.log 2>&1; echo rc=$? >> $O/r2c_pytest.log
timeout 300 python scripts/unet_profile.py 32 5 > $O/r2c_unet.log 2>&1
LFM_UNET_FUSE_OUT=0 timeout 300 python scripts/unet_profile.py 32 5 >> $O/r2c_unet.log 2>&1
timeout 300 python scripts/edm_profile.py >> $O/r2c_unet.log 2>&1
timeout 300 python scripts/vae_profile.py 1 16 > $O/r2c_vae.log 2>&1
LFM_UNET_FUSE_OUT=0 timeout 300 python scripts/vae_profile.py 16 >> $O/r2c_vae.log 2>&1
LFM_BN_QKV=640 LFM_BN_PROJ=640 LFM_BN_FC1=640 LFM_BN_FC2=640 LFM_BENCH_GEMM_BN=640 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2c_bench640.json 2> $O/r2c_bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/r2c_bench512.json 2>> $O/r2c_bench.err
tail -4 $O/r2c_gemm_test.log; grep -v "^===" $O/r2c_perf.log | tail -34; tail -12 $O/r2c_pytest.log; cat $O/r2c_unet.log $O/r2c_vae.log; cut -c1-250 $O/r2c_bench640.json; cut -c1-250 $O/r2c_bench512.json
