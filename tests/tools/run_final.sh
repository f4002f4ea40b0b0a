# Round-end evidence run (one GPU): regression, smoke, bench line, ncu launch lists, other-config records.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
timeout 600 python bench.py 2>gpurun_out/final_bench.err | tail -1 > gpurun_out/final_bench.json; cat gpurun_out/final_bench.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 600 -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/final_ncu_bench.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/final_unet_launches.csv python scripts/unet_profile.py 32 1 > gpurun_out/final_ncu_unet.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/final_edm_launches.csv python scripts/edm_profile.py 64 1 > gpurun_out/final_ncu_edm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name attention3_t256_d64 --launch-skip 30 -c 1 -o gpurun_out/final_attention3 python scripts/dit_profile.py 64 1 > gpurun_out/final_ncu_attn.log 2>&1
timeout 600 python scripts/bench_configs.py 2>&1 | grep "^{" > gpurun_out/final_other_configs.jsonl; cat gpurun_out/final_other_configs.jsonl
