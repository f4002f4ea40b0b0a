"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): mini DiT forward + Euler/Heun, mini UNet forward."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lfm_b200  # noqa: E402
from lfm_b200.synthetic import synthetic_state_dict, synthetic_unet_state_dict  # noqa: E402

dev = torch.device("cuda:0")
net = lfm_b200.DiT(img_resolution=32, patch_size=2, in_channels=4, hidden_size=256, depth=2, num_heads=4, label_dropout=0.1, num_classes=10)
net.load_state_dict(synthetic_state_dict(net, 3), strict=True)
net = net.to(dev)
x = torch.randn(3, 4, 32, 32, device=dev)
y = torch.tensor([1, 2, 3], device=dev)
v = net(torch.tensor([0.3, 0.5, 0.7], device=dev), x, y)
ynull = torch.full((3,), 10, device=dev)
out = lfm_b200.karras_sample(net, torch.cat([x, x]), 3, clip_denoised=False, model_kwargs=dict(y=torch.cat([y, ynull]), cfg_scale=1.5),
                             sigma_min=1e-5, sigma_max=1.0, sampler="heun")
un = lfm_b200.UNetModel(image_size=32, in_channels=4, model_channels=128, out_channels=4, num_res_blocks=1, attention_resolutions=(2, 4),
                        channel_mult=(1, 2, 2), num_heads=2, use_scale_shift_norm=True)
un.load_state_dict(synthetic_unet_state_dict(un, 5), strict=True)
un = un.to(dev)
vu = un(torch.tensor(0.4, device=dev), x)
torch.cuda.synchronize()
print("ok", float(v.abs().mean()), float(out.abs().mean()), float(vu.abs().mean()))
