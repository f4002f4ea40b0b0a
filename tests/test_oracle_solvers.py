"""Known-answer tests for the torchdiffeq restatements (euler fixed grid, dopri5) in oracle/solvers.

torchdiffeq is not in the image and the reference pins no results for it, so these anchor the
restatement on closed forms, NFE counts and scipy's independent Dormand-Prince implementation.
"""
import math

import pytest

import numpy as np
import torch

from oracle import solvers as osol


def test_euler_grid_nodes_and_nfe():
    for n in (10, 20, 50, 100):
        tk = osol.tdq_euler_grid(1.0 / n)
        assert len(tk) == n + 1 and float(tk[0]) == 1.0 and float(tk[-1]) == 0.0
        assert abs(float(tk[1]) - (1 - 1.0 / n)) < 1e-6
        assert float(tk[-2]) > 0  # the model never sees t = 0
    # default --step_size 0.01 => 100 NFE
    assert len(osol.tdq_euler_grid(0.01)) - 1 == 100


def test_euler_linear_field_closed_form():
    # dx/dt = a x integrated from t=1 down to 0 with N Euler steps: x_N = x_0 (1 - a/N)^N
    a, n = 0.7, 20
    x0 = torch.randn(3, 4, 8, 8, generator=torch.Generator().manual_seed(0))
    xf, nfe = osol.tdq_euler(lambda t, x: a * x, x0, 1.0 / n)
    assert nfe == n
    assert torch.allclose(xf, x0 * (1 - a / n) ** n, rtol=2e-5, atol=1e-6)


def test_karras_euler_heun_closed_form():
    a = -0.5
    x0 = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(1))
    sig = osol.karras_sigmas(9)
    calls = []

    def f(t, x):
        calls.append(t.clone())
        return a * x

    xe = osol.sample_euler(f, x0, sig)
    assert len(calls) == 8 and calls[0].shape == (2,)  # NFE = steps - 1, t as [B] vector
    ref = x0.clone()
    for i in range(8):
        ref = ref * (1 + a * float(sig[i + 1] - sig[i]))
    assert torch.allclose(xe, ref, rtol=1e-5, atol=1e-6)
    calls.clear()
    xh = osol.sample_heun(f, x0, sig)
    assert len(calls) == 16
    ref = x0.clone()
    for i in range(8):
        h = float(sig[i + 1] - sig[i])
        ref = ref * (1 + a * h + 0.5 * (a * h) ** 2)
    assert torch.allclose(xh, ref, rtol=1e-5, atol=1e-6)


def test_heun_nfe_quirk_counts():
    n = [0]

    def f(t, x):
        n[0] += 1
        return -x

    x0 = torch.ones(1, 1, 2, 2)
    osol.karras_sample(f, x0, 50, "heun")
    assert n[0] == 88  # intervals 39..48 are Euler-only
    n[0] = 0
    osol.karras_sample(f, x0, 25, "heun")
    assert n[0] == 48


def test_dopri5_tableau_identities():
    for a, b in zip(osol._DP_ALPHA, osol._DP_BETA):
        assert abs(sum(b) - a) < 1e-14
    assert abs(sum(osol._DP_CSOL) - 1) < 1e-14
    assert abs(sum(osol._DP_CERR)) < 1e-14
    assert abs(sum(osol._DP_MID) - 0.5) < 1e-12
    assert osol._DP_CSOL[:6] == osol._DP_BETA[-1]


def test_dopri5_linear_field_and_nfe():
    a = 1.3
    x0 = torch.randn(4, 4, 8, 8, generator=torch.Generator().manual_seed(2))
    xf, st = osol.tdq_dopri5(lambda t, x: a * x, x0, rtol=1e-5, atol=1e-5)
    # integrating dx/dt = a x from t=1 to t=0 gives x0 * exp(-a)
    assert torch.allclose(xf, x0 * math.exp(-a), rtol=2e-4, atol=2e-5)
    assert st.nfe == 2 + 6 * (st.accepted + st.rejected)
    assert st.accepted >= 2


def test_dopri5_vs_scipy_rk45_nonlinear():
    from scipy.integrate import solve_ivp

    x0 = torch.tensor([[0.3, -1.2, 0.8, 2.0]]).reshape(1, 4, 1, 1)

    def f(t, x):
        return torch.sin(3 * t) * x + torch.cos(x) + t

    xf, st = osol.tdq_dopri5(f, x0, rtol=1e-5, atol=1e-5)
    sol = solve_ivp(lambda t, y: np.sin(3 * t) * y + np.cos(y) + t, (1.0, 0.0), x0.reshape(-1).double().numpy(),
                    method="RK45", rtol=1e-10, atol=1e-12)
    assert np.allclose(xf.reshape(-1).numpy(), sol.y[:, -1], rtol=2e-4, atol=2e-4)
    assert st.nfe == 2 + 6 * (st.accepted + st.rejected)


def test_fixed_rk_orders_on_linear_field():
    # dx/dt = a x from t=1 to 0: exact x0 exp(-a); midpoint is 2nd order, rk4 (3/8 rule) 4th order
    a = 0.9
    x0 = torch.randn(2, 4, 4, 4, generator=torch.Generator().manual_seed(3)).double().float()
    exact = x0 * math.exp(-a)
    errs = {}
    for m, nfe_per in (("midpoint", 2), ("rk4", 4)):
        e = []
        for n in (5, 10):
            xf, nfe = osol.tdq_fixed_rk(lambda t, x: a * x, x0, 1.0 / n, m)
            assert nfe == nfe_per * n
            e.append(float((xf - exact).abs().max()))
        errs[m] = e
    assert errs["midpoint"][0] / errs["midpoint"][1] > 3.0       # ~4x per halving
    assert errs["rk4"][0] < 1e-4 and errs["rk4"][1] < 2e-5
    # closed forms of one step: midpoint 1 + z + z^2/2, rk4 1 + z + z^2/2 + z^3/6 + z^4/24 with z = -a h
    xf, _ = osol.tdq_fixed_rk(lambda t, x: a * x, x0, 1.0, "rk4")
    z = -a
    assert torch.allclose(xf, x0 * (1 + z + z * z / 2 + z ** 3 / 6 + z ** 4 / 24), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["rot", "dit"])
def test_tdq_fixtures_from_real_torchdiffeq(name):
    """S2 / S3 pin: the restated torchdiffeq euler / midpoint / rk4 / dopri5 against fixtures recorded from the REAL
    package by oracle/make_tdq_goldens.py.  torchdiffeq is not in the build image, so the fixtures do not exist yet
    and the test skips - parity for these two rows stays 'unpinned'  until they do."""
    import os
    import sys
    import numpy as np
    from tests._util import GOLDEN, rel_l2
    path = os.path.join(GOLDEN, f"tdq_{name}.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/tdq_*.npz not recorded: torchdiffeq is absent from the image (oracle/make_tdq_goldens.py)")
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "oracle"))
    from oracle.make_tdq_goldens import fields
    z = np.load(path)
    f, x0, _ = fields()[name]
    assert np.array_equal(x0.numpy(), z["x0"])
    for h in (0.25, 0.1, 0.02, 1.0 / 3.0):
        key = f"euler_h{h:.6f}"
        out, nfe = osol.tdq_euler(f, x0, h)
        assert nfe == int(z[key + "_nfe"])
        assert np.array_equal(osol.tdq_euler_grid(h)[:-1].numpy(), z[key + "_times"])      # the times the model sees
        assert rel_l2(out, z[key]) < 1e-6
    for m, hh in (("midpoint", 0.2), ("rk4", 0.2)):
        out, nfe = osol.tdq_fixed_rk(f, x0, hh, m)
        assert nfe == int(z[f"{m}_h{hh}_nfe"]) and rel_l2(out, z[f"{m}_h{hh}"]) < 1e-6
    for tol in (1e-2, 1e-3, 1e-5):
        key = f"dopri5_tol{tol:g}"
        out, st = osol.tdq_dopri5(f, x0, rtol=tol, atol=tol)
        assert st.nfe == int(z[key + "_nfe"])                                                # same accept/reject sequence
        assert rel_l2(out, z[key]) < 1e-5


def test_other_adaptive_pairs_known_answers():
    """bosh3 / adaptive_heun restatements (torchdiffeq bosh3.py / adaptive_heun.py from memory; unpinned): tableau identities,
    NFE = 2 + stages * steps, convergence on dx/dt = -x, and the Bogacki-Shampine weights against scipy's RK23 tableau."""
    from scipy.integrate import RK23
    tb = osol._TABLEAUS["bosh3"]
    assert np.allclose(tb["c_sol"][:3], RK23.B) and np.allclose(tb["alpha"][:2], RK23.C[1:]) and np.allclose(-RK23.E, tb["c_err"])      # scipy stores low - high
    for name, stages in (("bosh3", 3), ("adaptive_heun", 1), ("dopri5", 6)):
        t = osol._TABLEAUS[name]
        assert abs(sum(t["c_sol"]) - 1) < 1e-12 and abs(sum(t["c_err"])) < 1e-12 and len(t["alpha"]) == stages
        for a, b in zip(t["alpha"], t["beta"]):
            assert abs(sum(b) - a) < 1e-12
        x0 = torch.tensor([[1.0, -2.0, 0.5]])
        errs = []
        for tol in (1e-3, 1e-5):
            out, st = osol.tdq_adaptive(lambda tt, xx: -xx, x0, name, tol, tol)
            assert st.nfe == 2 + stages * (st.accepted + st.rejected)
            errs.append(float((out / x0 - math.e).abs().max()))
        assert errs[1] < errs[0] and errs[1] < 5e-3

