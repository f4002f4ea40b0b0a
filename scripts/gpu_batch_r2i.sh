#!/bin/bash
# 2-GPU validation of the bench (extras at world > 1) + reference arm
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2i_bench_n2.json 2> $O/r2i_bench_n2.err
echo rc=$? >> $O/r2i_bench_n2.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/r2i_bench_ref.json 2> $O/r2i_bench_ref.err
cut -c1-1500 $O/r2i_bench_n2.json; tail -5 $O/r2i_bench_n2.err; cut -c1-600 $O/r2i_bench_ref.json
