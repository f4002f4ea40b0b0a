#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I lfm_b200/csrc tests/tools/l2_residency.cu -o /tmp/l2res > $O/r2t_build.log 2>&1
timeout 120 /tmp/l2res > $O/r2t_l2res.log 2>&1; echo rc=$? >> $O/r2t_l2res.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "fixture or properties or odd_batch or cfg_identity or smoke" --timeout=600 > $O/r2t_pytest.log 2>&1; echo rc=$? >> $O/r2t_pytest.log
timeout 200 python scripts/dit_profile.py 64 20 > $O/r2t_dit.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/r2t_bench.json 2> $O/r2t_bench.err
cat $O/r2t_l2res.log; tail -4 $O/r2t_pytest.log; cat $O/r2t_dit.log; cut -c1-200 $O/r2t_bench.json
