"""Parameter-estimation support (BASELINE config 5): finite-difference sensitivities of the N1 outlet with respect to
kinetic constants, evaluated as ONE ensemble launch.

The reference has no estimation code ("parameter estimation" appears only as prose in `README.md:8`); what an
estimator needs from this path is the outlet vector of many reactors under perturbed kinetic constants.  The constants are
scalar entries of the ``VARS`` dict promoted to per-instance runtime parameters (``params=``), so the base case and all
perturbed variants of all reactors run through the same compiled kernel without recompiling.
"""
from __future__ import annotations

import numpy as np


def perturbation_table(model, modelInput, sweep, theta, rel=1e-4):
    """inst_in rows for N base reactors followed by len(theta) blocks of N rows with theta_k*(1+rel).

    ``theta``: ordered names of scalar VARS entries (must be the model's ``params``).  Returns (table, steps[k])."""
    base = model.instance_table(modelInput, sweep)
    N = base.shape[0]
    cols = [9 + model.n + model.params.index(name) for name in theta]
    blocks, steps = [base], []
    for c in cols:
        t = base.copy()
        h = t[:, c]*rel
        t[:, c] += h
        blocks.append(t)
        steps.append(h)
    return np.concatenate(blocks, axis=0), np.stack(steps, axis=1), N


def outlet_jacobian_fd(modelInput, sweep, theta, rel=1e-4, rtol=1e-11, atol=1e-15, device=0):
    """d(outlet)/d(theta_k) [N, nv1, p] by forward differences; solver tolerance tight enough that integration
    noise stays below the perturbation.  Also returns the base outlets [N, nv1]."""
    from .api import get_model
    model = get_model(modelInput, params=tuple(theta), device=device)
    tab, steps, N = perturbation_table(model, modelInput, sweep, theta, rel)
    r = model.solve_n1_host(tab, rtol=rtol, atol=atol, n_out=2, profile=False)
    if (r["status"] != 0).any():
        raise RuntimeError(f"{int((r['status'] != 0).sum())} of {len(tab)} perturbed solves failed")
    out = r["outlet"].reshape(1 + len(theta), N, model.nv1)
    J = (out[1:] - out[0][None])/np.moveaxis(steps, 1, 0)[:, :, None]
    return np.moveaxis(J, 0, 2), out[0]


def outlet_jacobian_sens(modelInput, sweep, theta, rtol=1e-9, atol=1e-13, device=0):
    """d(outlet)/d(theta_k) [N, nv1, p] by FORWARD SENSITIVITIES: the rate compiler differentiates the rate
    expressions with respect to the runtime parameters, and K2s integrates s = dx/dtheta_k next to the state with the same
    factorised Newton matrix.  One launch of N x p instances at the working tolerance -- no perturbation size to choose, no
    subtraction of nearly equal outlets (forward differences need 7 solves per reactor at rtol 1e-11 and carry an O(rel) truncation
    error).  Also returns the base outlets [N, nv1]."""
    import torch
    from .api import get_model
    model = get_model(modelInput, params=tuple(theta), device=device)
    base = model.instance_table(modelInput, sweep)
    N, p = base.shape[0], len(theta)
    tab = torch.from_numpy(np.concatenate([base]*p, axis=0)).to(model._dev())          # parameter-major: a warp shares kp
    kp = torch.arange(p, dtype=torch.int32, device=tab.device).repeat_interleave(N)
    par = model.prolog(tab)
    r = model.solve_n1_sens(par, kp, rtol=rtol, atol=atol)
    bad = int((r["status"] != 0).sum().item())
    if bad:
        raise RuntimeError(f"{bad} of {N*p} sensitivity solves failed")
    J = r["doutlet"].view(p, N, model.nv1).permute(1, 2, 0).contiguous().cpu().numpy()
    return J, r["outlet"][:N].cpu().numpy()
