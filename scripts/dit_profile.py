"""DiT-L/2 forward, batch 64 (the headline workload's network evaluation) - used under ncu and for timing.
usage: dit_profile.py [batch] [iters] [model_type] [latent side]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lfm_b200  # noqa: E402
from lfm_b200.synthetic import synthetic_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
mt = sys.argv[3] if len(sys.argv) > 3 else "DiT-L/2"
S = int(sys.argv[4]) if len(sys.argv) > 4 else 32
dev = torch.device("cuda:0")
with torch.device("meta"):
    net = lfm_b200.DiT_models[mt](img_resolution=S, in_channels=4, label_dropout=0.0, num_classes=1)
# algorithmic FLOPs per sample per evaluation: 161,386,856,448 for DiT-L/2, 46,003,912,704 for DiT-B/2
L_, D_, P_ = net.depth, net.hidden_size, net.patch_size ** 2 * net.in_channels
T_ = (net.img_resolution // net.patch_size) ** 2
FLOPS = L_ * (24 * T_ * D_ * D_ + 4 * T_ * T_ * D_ + 12 * D_ * D_) + 2 * T_ * P_ * D_ + 2 * (256 * D_ + D_ * D_) + 4 * D_ * D_ + 2 * T_ * D_ * P_
sd = synthetic_state_dict(net, 1)
net = net.to_empty(device="cpu")
net.load_state_dict(sd, strict=True)
net = net.to(dev)
x = torch.randn(B, 4, S, S, device=dev)
t = torch.tensor(0.5, device=dev)
for _ in range(3):
    net(t, x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    v = net(t, x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"{mt} ({S}x{S} latents) forward B={B}: {ms:.3f} ms/NFE, {B * FLOPS / ms / 1e9:.1f} TFLOP/s, launches/NFE {net.launch_count() // (iters + 3)}")
