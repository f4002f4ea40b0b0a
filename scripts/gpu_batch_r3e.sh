#!/bin/bash
# round 2, session 2, final evidence run (one GPU): whole GPU suite, smoke(), the full bench line, the ncu launch list of one bench step and
# ncu --set full captures of the three kernels of a DiT block (pair GEMM on its fc1 instance, LayerNorm pass, attention).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu > $O/r3e_pytest.log 2>&1; echo rc=$? >> $O/r3e_pytest.log; tail -n 3 $O/r3e_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | tee $O/r3e_smoke.log
timeout 900 python bench.py 2> $O/r3e_bench_full.err | tail -n 1 > $O/r3e_bench_full.json; cut -c1-300 $O/r3e_bench_full.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 600 -c 400 --csv --log-file $O/r3e_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/r3e_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm2_bf16_tcgen05|ln_modulate_kernel|attention3" --launch-skip 50 -c 7 -o $O/r3e_block_full python scripts/dit_profile.py 64 1 > $O/r3e_ncu_full.log 2>&1
ls -la $O/r3e_block_full.ncu-rep
echo finished
