"""DhariwalUNet (EDM-style ADM, ffhq_adm / imnet_adm preset: nf 256, ch_mult 1 2 3 4, attention at 16 / 8 / 4, 32x32x4
latents) forward - used under ncu and for timing.  usage: edm_profile.py [batch] [iters] [label_dim]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lfm_b200  # noqa: E402
from lfm_b200.synthetic import synthetic_edm_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
label_dim = int(sys.argv[3]) if len(sys.argv) > 3 else 0
FLOPS = 73.05e9   # per sample per NFE (convolutions, linears, attention matmuls; tests/test_oracle_edm.py)
dev = torch.device("cuda:0")
with torch.device("meta"):
    net = lfm_b200.DhariwalUNet(img_resolution=32, in_channels=4, out_channels=4, label_dim=label_dim, model_channels=256,
                                channel_mult=(1, 2, 3, 4), num_blocks=2, attn_resolutions=(16, 8, 4))
sd = synthetic_edm_state_dict(net, 1)
net = net.to_empty(device="cpu")
net.load_state_dict(sd, strict=True)
net = net.to(dev)
x = torch.randn(B, 4, 32, 32, device=dev)
y = torch.randint(0, label_dim, (B,), device=dev) if label_dim else None
t = torch.tensor(0.5, device=dev)
for _ in range(2):
    net(t, x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    v = net(t, x, y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"DhariwalUNet (adm preset) forward B={B}: {ms:.2f} ms/NFE, {B * FLOPS / ms / 1e9:.1f} TFLOP/s, "
      f"launches/NFE {net.launch_count() // (iters + 2)}")
