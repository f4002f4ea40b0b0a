#!/usr/bin/env python
"""Drop-in name for the reference's torchrun sampling CLI (test_flow_latent_ddp.py):
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 test_flow_latent_ddp.py ... --compute_fid
One process per GPU, per-rank batches, seed + rank, file index j * world + rank + total."""
from lfm_b200.cli import main

if __name__ == "__main__":
    raise SystemExit(main())
