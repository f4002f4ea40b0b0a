"""Timing of the native VAE decode (lfm_b200.AutoencoderKL, synthetic weights) on cuda:0: ms per batch, images/s, TFLOP/s.
python scripts/vae_profile.py [batch ...]   (wrap in ncu for a launch list: the decode of one batch is ~200 launches)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lfm_b200  # noqa: E402

dev = torch.device("cuda:0")
vae = lfm_b200.AutoencoderKL(max_batch=16)
vae.load_state_dict(lfm_b200.synthetic_vae_state_dict(vae, 1), strict=True)
vae = vae.to(dev).eval()
fl = vae.decode_flops_per_image(32)
for B in [int(a) for a in sys.argv[1:]] or [1, 16, 64]:
    z = torch.randn(B, 4, 32, 32, device=dev) * 1.1
    for _ in range(2):
        u8 = vae.decode_to_uint8(z)
    torch.cuda.synchronize()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = vae.launch_count()
    e0.record()
    for _ in range(reps):
        u8 = vae.decode_to_uint8(z)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"vae_decode_batch": B, "ms": round(ms, 3), "ms_per_image": round(ms / B, 3), "images_per_s": round(B / ms * 1e3, 1),
                      "tflops": round(B * fl / ms / 1e9, 1), "launches_per_decode": (vae.launch_count() - l0) // reps,
                      "finite": bool(torch.isfinite(vae.decode(z[:1]).sample).all())}), flush=True)
