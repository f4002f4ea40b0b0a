set -x
for p in 0 1 0 1; do LFM_PDL=$p timeout 120 python scripts/dit_profile.py 64 40 2>&1 | tail -1; done
LFM_PDL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "fixture or full_size or solvers or sampler or euler or heun or dopri" 2>&1 | tail -3
LFM_PDL=1 timeout 200 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('PDL=1 bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
LFM_PDL=0 timeout 200 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('PDL=0 bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
