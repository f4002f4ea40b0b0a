"""Oracle: ODE solvers of the sampling path (TEST INFRASTRUCTURE, see oracle/__init__).

``f(t, x) -> v`` is any callable with the reference's model signature (t first; 0-d or [B]).

Restated from
* ``sampler/karras_sample.py:7-76``   ``karras_sample`` (sigma grid, CFG-aware denoiser)
* ``sampler/karras_sample.py:86-118`` ``sample_euler``
* ``sampler/karras_sample.py:122-161````sample_heun`` (corrector guard ``i < steps - 1`` with the
  function default ``steps=40`` that ``karras_sample`` never overrides: SURVEY.md App. B.3)
* ``test_flow_latent.py:42-76``       ``sample_from_model`` -> torchdiffeq ``odeint_adjoint``.
  torchdiffeq is a third-party dependency (requirements.txt:3, unpinned, not in /root/reference).
  Its published algorithm (v0.2.3: ``misc._check_inputs`` time reversal + ``_PerturbFunc``,
  ``fixed_grid.Euler``, ``rk_common.RKAdaptiveStepsizeODESolver`` with the ``dopri5`` tableau) is
  restated below; parity for these two functions is UNPINNED by the reference (no tests there) and
  is anchored on closed-form / scipy known-answer tests in ``tests/test_oracle_solvers.py``.
"""
from __future__ import annotations

import math

import torch

# ----------------------------------------------------------------------------------------------
# Karras-style fixed-step samplers (flow matching: the "denoiser" output is the velocity itself)


def karras_sigmas(steps: int, sigma_min=1e-5, sigma_max=1.0) -> torch.Tensor:
    """karras_sample.py:30 with the values passed at test_flow_latent.py:87-88."""
    return torch.linspace(sigma_max, sigma_min, steps)


@torch.no_grad()
def sample_euler(f, x, sigmas):
    """karras_sample.py:86-118.  t is passed as a [B] vector; NFE = len(sigmas) - 1."""
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        v = f(sigmas[i] * s_in, x)
        x = x + v * (sigmas[i + 1] - sigmas[i])
    return x


@torch.no_grad()
def sample_heun(f, x, sigmas, corrector_limit: int = 39):
    """karras_sample.py:122-161 with s_churn = 0 (gamma = 0 => x_hat == x_cur exactly; the noise
    draw is multiplied by 0).  ``corrector_limit`` = ``steps - 1`` of the reference's guard
    ``if i < steps - 1`` where ``steps`` is the *function default 40* => intervals 39.. are
    Euler-only.  Pass ``len(sigmas)`` (or more) for a corrector on every interval."""
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        t_cur, t_next = sigmas[i], sigmas[i + 1]
        d_cur = f(t_cur * s_in, x)
        x_next = x + (t_next - t_cur) * d_cur
        if i < corrector_limit:
            d_prime = f(t_next * s_in, x_next)
            x_next = x + (t_next - t_cur) * (0.5 * d_cur + 0.5 * d_prime)
        x = x_next
    return x


def karras_sample(f, x_T, steps, sampler="heun", sigma_min=1e-5, sigma_max=1.0, corrector_limit=39):
    sig = karras_sigmas(steps, sigma_min, sigma_max)
    if sampler == "euler":
        return sample_euler(f, x_T, sig)
    if sampler == "heun":
        return sample_heun(f, x_T, sig, corrector_limit)
    raise KeyError(sampler)


# ----------------------------------------------------------------------------------------------
# torchdiffeq fixed-grid Euler over t = [1, 0]


def tdq_euler_grid(step_size: float, t0: float = 1.0, t1: float = 0.0) -> torch.Tensor:
    """The time nodes the MODEL sees (real, decreasing time) for torchdiffeq ``method='euler'``,
    ``options={'step_size': h}`` over ``t=[t0, t1]`` with t0 > t1.

    torchdiffeq negates decreasing time (s = -t) and builds
    ``s_k = arange(niters) * h + s_0`` (fp32), ``niters = ceil((s_end - s_0)/h + 1)``, then forces
    the last node to s_end.  Returns the fp32 nodes ``t_k = -s_k`` INCLUDING the final node, so
    step k evaluates f(t_k) and advances by ``dt_k = s_{k+1} - s_k``; NFE = len - 1.
    """
    s = -torch.tensor([t0, t1], dtype=torch.float32)
    niters = int(torch.ceil((s[-1] - s[0]) / step_size + 1).item())
    grid = torch.arange(0, niters, dtype=torch.float32) * step_size + s[0]
    grid[-1] = s[-1]
    return -grid


@torch.no_grad()
def tdq_euler(f, x0, step_size: float, t0: float = 1.0, t1: float = 0.0, perturb: bool = False):
    """``odeint(f, x0, [t0, t1], method='euler', options=dict(step_size=h))[-1]``.

    In negated time the wrapped field is ``-f(-s, y)`` and ``y += (s_{k+1} - s_k) * (-f)``; the
    model receives a 0-d fp32 t.  The last step lands exactly on s_end so no interpolation occurs.
    Returns (x_final, nfe).
    """
    tk = tdq_euler_grid(step_size, t0, t1)
    sk = -tk
    y = x0
    for k in range(len(tk) - 1):
        dt = sk[k + 1] - sk[k]
        # options["perturb"]: Perturb.NEXT on the step's evaluation = nextafter(s_k, s_k + 1) in negated time
        tm = -torch.nextafter(sk[k], sk[k] + 1) if perturb else tk[k]
        y = y + dt * (-f(tm, y))
    return y, len(tk) - 1


@torch.no_grad()
def tdq_fixed_rk(f, x0, step_size: float, method: str, t0: float = 1.0, t1: float = 0.0, perturb: bool = False):
    """torchdiffeq fixed-grid ``midpoint`` / ``rk4`` (v0.2.3 fixed_grid.py; ``rk4`` is the 3/8-rule
    ``rk4_alt_step_func``) on the same grid as :func:`tdq_euler`.  Integrates s = -t with the field -f(-s, y);
    stage times and dt are fp32 tensors, ``_one_third`` / ``_two_thirds`` python floats.  Returns (x_final, nfe)."""
    tk = tdq_euler_grid(step_size, t0, t1)
    sk = -tk
    nfe = [0]

    def ft(s, y):  # wrapped field in negated time
        nfe[0] += 1
        return -f(-s, y)

    y = x0
    one_third, two_thirds = 1.0 / 3.0, 2.0 / 3.0
    for k in range(len(tk) - 1):
        s0, s1 = sk[k], sk[k + 1]
        dt = s1 - s0
        s0e = torch.nextafter(s0, s0 + 1) if perturb else s0      # Perturb.NEXT on the first evaluation of the step
        s1e = torch.nextafter(s1, s1 - 1) if perturb else s1      # Perturb.PREV on rk4's last evaluation
        if method == "midpoint":
            half_dt = 0.5 * dt
            f0 = ft(s0e, y)
            y_mid = y + f0 * half_dt
            y = y + dt * ft(s0 + half_dt, y_mid)
        elif method == "rk4":
            k1 = ft(s0e, y)
            k2 = ft(s0 + dt * one_third, y + dt * k1 * one_third)
            k3 = ft(s0 + dt * two_thirds, y + dt * (k2 - k1 * one_third))
            k4 = ft(s1e, y + dt * (k1 - k2 + k3))
            y = y + (k1 + 3 * (k2 + k3) + k4) * dt * 0.125
        else:
            raise KeyError(method)
    return y, nfe[0]


# ----------------------------------------------------------------------------------------------
# torchdiffeq dopri5

_DP_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
_DP_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
_DP_CSOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
_DP_CERR = [
    35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
    -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1.0 / 60.0,
]
_DP_MID = [
    6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
    187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2,
]


def _rms(x: torch.Tensor) -> torch.Tensor:
    return x.pow(2).mean().sqrt()  # over ALL elements of the batch tensor


class Dopri5Stats:
    def __init__(self):
        self.nfe = 0
        self.accepted = 0
        self.rejected = 0


# torchdiffeq's other Runge-Kutta pairs that the reference CLI accepts (test_flow_latent.py:27 ADAPTIVE_SOLVER; the same
# rk_common.RKAdaptiveStepsizeODESolver with another tableau).  Restated from memory of torchdiffeq 0.2.3 (bosh3.py,
# adaptive_heun.py); UNPINNED like dopri5, anchored on tableau identities and scipy's RK23 (= Bogacki-Shampine).
_TABLEAUS = {
    "dopri5": dict(order=5, alpha=_DP_ALPHA, beta=_DP_BETA, c_sol=_DP_CSOL, c_err=_DP_CERR, c_mid=_DP_MID),
    "bosh3": dict(order=3, alpha=[1 / 2, 3 / 4, 1.0], beta=[[1 / 2], [0.0, 3 / 4], [2 / 9, 1 / 3, 4 / 9]],
                  c_sol=[2 / 9, 1 / 3, 4 / 9, 0.0], c_err=[2 / 9 - 7 / 24, 1 / 3 - 1 / 4, 4 / 9 - 1 / 3, -1 / 8],
                  c_mid=[0.0, 0.5, 0.0, 0.0]),
    "adaptive_heun": dict(order=2, alpha=[1.0], beta=[[1.0]], c_sol=[0.5, 0.5], c_err=[0.5, -0.5], c_mid=[0.5, 0.0]),
}


def tdq_dopri5(f, x0, rtol=1e-5, atol=1e-5, t0: float = 1.0, t1: float = 0.0, max_steps=100000):
    return tdq_adaptive(f, x0, "dopri5", rtol, atol, t0, t1, max_steps)


@torch.no_grad()
def tdq_adaptive(f, x0, method="dopri5", rtol=1e-5, atol=1e-5, t0: float = 1.0, t1: float = 0.0, max_steps=100000):
    """``odeint(f, x0, [t0, t1], method=method, rtol, atol, options=dict(dtype=float64))[-1]`` for the Runge-Kutta pairs of
    ``_TABLEAUS`` (rk_common.RKAdaptiveStepsizeODESolver).

    State y in fp32; time, dt and the controller in fp64; stage times/dt are cast to fp32 inside the
    Runge-Kutta step; the model sees a 0-d fp32 t; stages with alpha == 1 are evaluated one fp32
    ulp before the step end (``Perturb.PREV``).  The error norm is the RMS over the whole batch
    tensor, so the step sequence depends on the batch.  ``_select_initial_step`` runs with ``order - 1``, the step
    controller with ``order``; ``y1`` is the last stage's input when the tableau is first-same-as-last, else
    ``y0 + dt * sum c_sol k``; ``f1`` is ALWAYS the last stage derivative.  Returns (x_final, Dopri5Stats).
    """
    tab = _TABLEAUS[method]
    order = tab["order"]
    nst = len(tab["alpha"])
    st = Dopri5Stats()
    f32, f64 = torch.float32, torch.float64

    def func(s, y, perturb=0):
        # _PerturbFunc(_ReverseFunc(f)): cast s to y.dtype, nudge one ulp, evaluate -f(-s, y)
        s = torch.as_tensor(s).to(y.dtype)
        if perturb < 0:
            s = torch.nextafter(s, s - 1)
        st.nfe += 1
        return -f(-s, y)

    rtol_t = torch.tensor(rtol, dtype=f64)
    atol_t = torch.tensor(atol, dtype=f64)
    s0 = torch.tensor(-t0, dtype=f64)
    s_end = torch.tensor(-t1, dtype=f64)
    y0 = x0.to(f32)

    # _before_integrate + _select_initial_step(order - 1)
    f0 = func(s0, y0)
    scale = (atol_t + y0.abs() * rtol_t).to(f32)
    d0 = _rms(y0 / scale)
    d1 = _rms(f0 / scale)
    if d0 < 1e-5 or d1 < 1e-5:
        h0 = torch.tensor(1e-6, dtype=f32)
    else:
        h0 = 0.01 * d0 / d1
    h0 = h0.abs()
    f1 = func(s0 + h0, y0 + h0 * f0)
    d2 = (_rms((f1 - f0) / scale) / h0).abs()
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = torch.max(torch.tensor(1e-6, dtype=f32), h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / float(order))
    dt = torch.min(100 * h0, h1.abs()).to(f64)

    alpha = torch.tensor(tab["alpha"], dtype=f32)
    beta = [torch.tensor(b, dtype=f32) for b in tab["beta"]]
    c_sol = torch.tensor(tab["c_sol"], dtype=f32)
    c_err = torch.tensor(tab["c_err"], dtype=f32)
    c_mid = torch.tensor(tab["c_mid"], dtype=f32)
    fsal = float(c_sol[-1]) == 0.0 and bool((c_sol[:-1] == beta[-1]).all())

    s_lo = s0           # rk_state.t0 of the last accepted step
    s_hi = s0           # rk_state.t1
    interp = None
    n = 0
    while s_end > s_hi:
        assert n < max_steps
        n += 1
        t0s = s_hi
        t1s = t0s + dt
        t0_32, dt_32, t1_32 = t0s.to(f32), dt.to(f32), t1s.to(f32)
        k = [f0]
        yi = y0
        for i in range(nst):
            if float(alpha[i]) == 1.0:
                ti, perturb = t1_32, -1
            else:
                ti, perturb = t0_32 + alpha[i] * dt_32, 0
            yi = y0 + torch.stack(k, dim=-1).matmul(beta[i] * dt_32).view_as(f0)
            k.append(func(ti, yi, perturb))
        K = torch.stack(k, dim=-1)
        if not fsal:
            yi = y0 + K.matmul(dt_32 * c_sol).view_as(f0)
        y1, f1 = yi, k[-1]
        err = K.matmul(dt_32 * c_err)
        tol = (atol_t + rtol_t * torch.max(y0.abs(), y1.abs())).to(f32)
        ratio = _rms(err / tol).abs()
        accept = bool(ratio <= 1)
        if accept:
            st.accepted += 1
            y_mid = y0 + K.matmul(dt_32 * c_mid).view_as(y0)
            fa, fb = k[0], k[-1]
            a = 2 * dt_32 * (fb - fa) - 8 * (y1 + y0) + 16 * y_mid
            b = dt_32 * (5 * fa - 3 * fb) + 18 * y0 + 14 * y1 - 32 * y_mid
            c = dt_32 * (fb - 4 * fa) - 11 * y0 - 5 * y1 + 16 * y_mid
            interp = [y0, dt_32 * fa, c, b, a]
            s_lo, s_hi = t0s, t1s
            y0, f0 = y1, f1
        else:
            st.rejected += 1
        # _optimal_step_size(order)
        if ratio == 0:
            dt = dt * 10.0
        else:
            r = ratio.to(f64)
            dfactor = 1.0 if ratio < 1 else 0.2
            factor = min(10.0, max(0.9 / float(r) ** (1.0 / float(order)), dfactor))
            dt = dt * factor
    # _interp_evaluate at s_end inside the last accepted step
    xq = ((s_end - s_lo) / (s_hi - s_lo)).to(f32)
    total = interp[0] + xq * interp[1]
    xp = xq
    for coeff in interp[2:]:
        xp = xp * xq
        total = total + xp * coeff
    return total, st
