"""ORACLE (test infrastructure, not product code) -- adaptive-step ODE solve of the `use_torchode=True` branch.

Reference call site: cfm_superresolution.py:259-276
    term = to.ODETerm(fn); step_method = to.Tsit5(term); controller = to.IntegralController(atol, rtol, term)
    solver = to.AutoDiffAdjoint(step_method, controller); sol = solver.solve(InitialValueProblem(y0, t_eval=t))
    sampled = sol.ys[:, -1]
torchode (pyproject.toml:15 `torchode==1.0.0`) is a third-party dependency that is NOT installed offline and not
vendored under /root/reference, so this file restates its published algorithm:
  * Tsit5: Tsitouras 2011 5(4) pair with FSAL (7 stages, 6 new field evaluations per step); Dopri5: Dormand-Prince.
  * IntegralController = PID controller with (p, i, d) = (0, 1, 0): error ratio r = rms(err / (atol + rtol*max(|y0|,|y1|))),
    accept when r < 1, dt_next = dt * clip(safety * r^(-1/k), factor_min, factor_max), safety 0.9, factors 0.2 / 10,
    k = convergence order of the method (5).
  * initial step: Hairer / Noersett / Wanner's two-evaluation heuristic (as torchdiffeq / torchode / scipy use it).
  * every problem instance of a batch has its own t, dt and accept / reject history -- the instances are independent,
    which is the point of torchode; here they are simply solved one after the other.
  * the step is clipped so the solver lands on t_end exactly, and only the final state is returned (the call site reads
    `sol.ys[:, -1]`; the interior t_eval points are never used).

PARITY UNPINNED against torchode itself: no torchode golden can be generated in this container.  The Dopri5 path IS pinned
against an installed independent implementation of the same method and controller family, scipy.integrate.solve_ivp('RK45')
(same step counts to +-2, same error; tests/test_next_rows_cpu.py).  What else the tests pin: the tableaux against
the Runge-Kutta order conditions, the solver against closed-form solutions at the requested tolerance, and (for the FLowHigh
field) against a 256-step fixed-grid solve of the already-pinned vector field.  Controller details that may differ from
torchode 1.0.0 (exact clipping constants, dense-output evaluation of the last point instead of step clipping) move the
result by O(tolerance), not more.
"""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import numpy as np
import torch

# c, a (lower-triangular rows), b (5th order), e = b - b_hat (error estimate weights)
TSIT5 = dict(
    order=5,
    c=[0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0],
    a=[[],
       [0.161],
       [-0.008480655492356989, 0.335480655492357],
       [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
       [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
       [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383],
       [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774]],
    e=[-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
       0.5823571654525552, -0.45808210592918697, 1.0 / 66.0],
)
DOPRI5 = dict(
    order=5,
    c=[0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0],
    a=[[],
       [1 / 5],
       [3 / 40, 9 / 40],
       [44 / 45, -56 / 15, 32 / 9],
       [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
       [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
       [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]],
    e=[71 / 57600, 0.0, -71 / 16695, 71 / 1920, -17253 / 339200, 22 / 525, -1 / 40],
)
TABLEAUX: Dict[str, dict] = {"tsit5": TSIT5, "dopri5": DOPRI5}

SAFETY, FACTOR_MIN, FACTOR_MAX = 0.9, 0.2, 10.0


def _rms(x: torch.Tensor) -> float:
    return float(torch.sqrt(torch.mean(x.double() ** 2)))


def initial_step(fn: Callable, t0: float, y0: torch.Tensor, f0: torch.Tensor, order: int, atol: float, rtol: float,
                 span: float) -> float:
    scale = atol + rtol * y0.abs()
    d0, d1 = _rms(y0 / scale), _rms(f0 / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    h0 = min(h0, span)
    f1 = fn(torch.tensor(t0 + h0, dtype=y0.dtype), y0 + h0 * f0)
    d2 = _rms((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / (order + 1))
    return min(100.0 * h0, h1, span)


def odeint_adaptive(fn: Callable, y0: torch.Tensor, t0: float, t1: float, *, atol: float = 1e-5, rtol: float = 1e-5,
                    method: str = "tsit5", max_steps: int = 10000) -> Tuple[torch.Tensor, dict]:
    """Final state y(t1) of one problem instance (y0 of any shape; fn(t, y) -> dy/dt).  Returns (y, stats)."""
    tab = TABLEAUX[method]
    a, c, e, k_ord = tab["a"], tab["c"], tab["e"], tab["order"]
    y, t = y0, float(t0)
    f = fn(torch.tensor(t, dtype=y0.dtype), y)
    nfe = 1
    dt = initial_step(fn, t, y, f, k_ord, atol, rtol, t1 - t0)
    nfe += 1
    n_steps = n_accept = 0
    while t < t1 and n_steps < max_steps:
        dt = min(dt, t1 - t)
        last = dt >= t1 - t
        ks = [f]
        for s in range(1, 7):
            ys = y
            for j, w in enumerate(a[s]):
                if w != 0.0:
                    ys = ys + (dt * w) * ks[j]
            ks.append(fn(torch.tensor(t + c[s] * dt, dtype=y0.dtype), ys))
            nfe += 1
        y1 = ys  # FSAL: stage 7 is evaluated AT the 5th-order solution (a[6] == b)
        err = sum((dt * w) * kk for w, kk in zip(e, ks) if w != 0.0)
        ratio = _rms(err / (atol + rtol * torch.maximum(y.abs(), y1.abs())))
        n_steps += 1
        if ratio < 1.0:
            t = t1 if last else t + dt
            y, f = y1, ks[6]
            n_accept += 1
        factor = FACTOR_MAX if ratio == 0.0 else min(FACTOR_MAX, max(FACTOR_MIN, SAFETY * ratio ** (-1.0 / k_ord)))
        dt = dt * factor
    if t < t1:
        raise RuntimeError(f"adaptive solve did not reach t_end in {max_steps} steps (t = {t})")
    return y, {"n_steps": n_steps, "n_accepted": n_accept, "n_f_evals": nfe}


def odeint_adaptive_batch(fn_b: Callable, y0: torch.Tensor, t0: float, t1: float, **kw):
    """Batch of independent instances: fn_b(b, t, y[1, ...]) is instance b's field."""
    outs, stats = [], []
    for b in range(y0.shape[0]):
        yb, st = odeint_adaptive(lambda t, y, b=b: fn_b(b, t, y), y0[b: b + 1], t0, t1, **kw)
        outs.append(yb)
        stats.append(st)
    return torch.cat(outs), stats


def order_condition_residuals(tab: dict) -> Dict[str, float]:
    """Residuals of the Runge-Kutta order conditions up to order 4 for b (the 5th-order weights = last row of a) and of
    the conditions up to order 3... for the embedded pair: sum(e) and sum(e c) must vanish (b and b_hat are both
    consistent, at least 2nd order); used by the CPU test."""
    n = 7
    A = np.zeros((n, n))
    for i, row in enumerate(tab["a"]):
        A[i, : len(row)] = row
    c = np.array(tab["c"])
    b = A[6].copy()
    e = np.array(tab["e"])
    one = np.ones(n)
    res = {
        "row_sums": float(np.abs(A @ one - c).max()),
        "b1": abs(b.sum() - 1), "b2": abs(b @ c - 1 / 2), "b3a": abs(b @ c ** 2 - 1 / 3), "b3b": abs(b @ (A @ c) - 1 / 6),
        "b4a": abs(b @ c ** 3 - 1 / 4), "b4b": abs((b * c) @ (A @ c) - 1 / 8), "b4c": abs(b @ (A @ c ** 2) - 1 / 12),
        "b4d": abs(b @ (A @ (A @ c)) - 1 / 24),
        "b5a": abs(b @ c ** 4 - 1 / 5), "b5b": abs(b @ (A @ c ** 3) - 1 / 20),
        "e0": abs(e.sum()), "e1": abs(e @ c), "e2a": abs(e @ c ** 2), "e2b": abs(e @ (A @ c)),
        "e3a": abs(e @ c ** 3), "e3d": abs(e @ (A @ (A @ c))),
    }
    return {k: float(v) for k, v in res.items()}
