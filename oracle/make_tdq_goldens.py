"""Record tests/golden/tdq_*.npz from the REAL torchdiffeq package - when it is importable.

Run:  python oracle/make_tdq_goldens.py      (anywhere torchdiffeq is installed; it is NOT in this image)

torchdiffeq is the third-party dependency behind ``test_flow_latent.py:61-73`` (``odeint_adjoint`` with
``method='euler'``/``'dopri5'``); it is absent from /root/reference and from the build image, so ``oracle.solvers``
restates its published algorithm and S2 / S3 parity is UNPINNED.  This script is what lifts
the cap the day the package is available: it drives ``torchdiffeq.odeint`` exactly as ``sample_from_model`` does
(t = [1, 0], fp32 state, ``options={"step_size": h, "perturb": False}`` for the fixed grid,
``atol = rtol``, ``options={"dtype": torch.float64}`` for dopri5) on (a) an analytic vector field and (b) the
oracle DiT ``mini_uncond`` network, and stores inputs, outputs and the NFE count.  ``tests/test_oracle_solvers.py::
test_tdq_fixtures_from_real_torchdiffeq`` then checks the restatement against these files (and skips while they do
not exist).  Without torchdiffeq the script says so and exits 0 without writing anything.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


class Counted(torch.nn.Module):
    def __init__(self, f):
        super().__init__()
        self.f, self.nfe, self.times = f, 0, []

    def forward(self, t, x):
        self.nfe += 1
        self.times.append(float(t))
        return self.f(t, x)


def fields():
    """name -> (f(t, x), x0).  `rot` is linear with a time-dependent rate (closed form exists); `dit` is the oracle
    network of the mini_uncond fixture (same weights as every other test)."""
    from oracle import dit as odit
    g = torch.Generator().manual_seed(123)
    x_rot = torch.randn(3, 4, 8, 8, generator=g)
    A = torch.randn(4, 4, generator=g) * 0.7
    rot = lambda t, x: torch.einsum("ij,bjhw->bihw", A, x) * (0.5 + t) - 0.3 * x  # noqa: E731
    cfg = odit.DiTConfig(img_resolution=32, patch_size=2, in_channels=4, hidden_size=256, depth=2, num_heads=4,
                         label_dropout=0.0, num_classes=1)
    sd = odit.synthetic_state_dict(cfg, 11)
    x_dit = torch.randn(2, 4, 32, 32, generator=g)
    dit = lambda t, x: odit.dit_forward(sd, cfg, t, x)  # noqa: E731
    return {"rot": (rot, x_rot, dict(A=A.numpy())), "dit": (dit, x_dit, {})}


def main():
    try:
        from torchdiffeq import odeint
        import torchdiffeq
    except ImportError:
        print("torchdiffeq is not installed here: nothing recorded (S2/S3 parity stays unpinned)")
        return 0
    t = torch.tensor([1.0, 0.0])
    for name, (f, x0, extra) in fields().items():
        rec = dict(x0=x0.numpy(), version=np.array(getattr(torchdiffeq, "__version__", "unknown")), **extra)
        with torch.no_grad():
            for h in (0.25, 0.1, 0.02, 1.0 / 3.0):
                m = Counted(f)
                out = odeint(m, x0, t, method="euler", atol=1e-5, rtol=1e-5, options={"step_size": h, "perturb": False})
                key = f"euler_h{h:.6f}"
                rec[key] = out[-1].numpy()
                rec[key + "_nfe"] = np.array(m.nfe)
                rec[key + "_times"] = np.array(m.times, dtype=np.float32)
            for m_name, hh in (("midpoint", 0.2), ("rk4", 0.2)):
                m = Counted(f)
                out = odeint(m, x0, t, method=m_name, options={"step_size": hh, "perturb": False})
                rec[f"{m_name}_h{hh}"] = out[-1].numpy()
                rec[f"{m_name}_h{hh}_nfe"] = np.array(m.nfe)
            for tol in (1e-2, 1e-3, 1e-5):
                m = Counted(f)
                out = odeint(m, x0, t, method="dopri5", atol=tol, rtol=tol, options={"dtype": torch.float64})
                key = f"dopri5_tol{tol:g}"
                rec[key] = out[-1].numpy()
                rec[key + "_nfe"] = np.array(m.nfe)
                rec[key + "_times"] = np.array(m.times, dtype=np.float32)
        path = os.path.join(OUT, f"tdq_{name}.npz")
        np.savez(path, **rec)
        print("wrote", path)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
