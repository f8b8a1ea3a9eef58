"""Independent pin for the oracle: the reference's PRIMAL program solved with scipy.  TEST INFRASTRUCTURE.

This is the closest thing to "running the reference" this image allows (cvxpy and every conic
solver are absent, no network).  It restates the optimisation problem exactly as the scripts
pose it to cvxpy -- variables Delta_i, Lambda_i >= 0 per pool (arbitrage.py:51-52), net trade
psi = sum_i A_i (Lambda_i - Delta_i) (arbitrage.py:54), new reserves R + gamma*D - L
(arbitrage.py:60), one trading-function constraint per pool (arbitrage.py:63-74) and the
utility constraints (arbitrage.py:77 | liquidation.py:77-80 | two-asset.py:86) -- and hands it to
a general NLP method (SLSQP, then a trust-constr polish), which shares no code and no
algorithmic idea with the dual decomposition in cfmm_oracle.py.  ``tests/golden/make_golden.py``
uses it to produce the committed golden vectors.
"""
from __future__ import annotations

import numpy as np
from scipy import optimize


def solve_primal(n_tokens, local_indices, reserves, fees, kinds, weights, objective, constraints,
                 x0=None, maxiter=2000):
    """objective: c (n-vector) -> maximise c'psi.
    constraints: list of ('ge'|'eq', j, a_j) meaning psi_j + a_j >= 0 | == 0."""
    m = len(local_indices)
    sizes = [len(l) for l in local_indices]
    off = np.concatenate([[0], np.cumsum(sizes)])
    nv = int(off[-1])
    R = [np.asarray(r, float) for r in reserves]
    W = []
    for i in range(m):
        if kinds[i] == "sum":
            W.append(None)
        elif kinds[i] == "bounded_product":
            W.append(np.asarray(weights[i], float))          # virtual offsets o, not weights
        else:
            w = np.ones(sizes[i]) if weights[i] is None else np.asarray(weights[i], float)
            W.append(w / w.sum())
    c = np.asarray(objective, float)
    scat = np.concatenate([np.asarray(l) for l in local_indices])

    def split(z):
        return z[:nv], z[nv:]

    def psi(z):
        D, L = split(z)
        out = np.zeros(n_tokens)
        np.add.at(out, scat, L - D)
        return out

    def newres(z, i):
        D, L = split(z)
        s = slice(off[i], off[i + 1])
        return R[i] + fees[i] * D[s] - L[s]

    cons = []
    for i in range(m):
        if kinds[i] == "sum":
            cons.append(dict(type="ineq", fun=lambda z, i=i: np.sum(newres(z, i)) - np.sum(R[i])))
            cons.append(dict(type="ineq", fun=lambda z, i=i: newres(z, i)))
        elif kinds[i] == "bounded_product":
            # sqrt((x0+o0)(x1+o1)) >= sqrt((R0+o0)(R1+o1)) in log form, and x >= 0
            cons.append(dict(type="ineq", fun=lambda z, i=i: np.sum(
                np.log(np.maximum(newres(z, i) + W[i], 1e-300)) - np.log(R[i] + W[i]))))
            cons.append(dict(type="ineq", fun=lambda z, i=i: newres(z, i)))
        else:
            # log form of geo_mean(x, p=w) >= geo_mean(R, p=w): same feasible set, better scaled
            cons.append(dict(type="ineq", fun=lambda z, i=i: np.dot(
                W[i], np.log(np.maximum(newres(z, i), 1e-300)) - np.log(R[i]))))
    for kind, j, aj in constraints:
        cons.append(dict(type="ineq" if kind == "ge" else "eq", fun=lambda z, j=j, aj=aj: psi(z)[j] + aj))
    f = lambda z: -float(np.dot(c, psi(z)))
    bounds = [(0, None)] * (2 * nv)
    best = None
    rng = np.random.default_rng(0)
    starts = [np.zeros(2 * nv)] if x0 is None else [np.asarray(x0, float)]
    starts += [0.1 * rng.random(2 * nv) for _ in range(3)]
    for z0 in starts:
        res = optimize.minimize(f, z0, method="SLSQP", bounds=bounds, constraints=cons,
                                options=dict(maxiter=maxiter, ftol=1e-15))
        feas = all((np.min(np.atleast_1d(cn["fun"](res.x))) >= -1e-9) if cn["type"] == "ineq"
                   else (abs(cn["fun"](res.x)) <= 1e-9) for cn in cons)
        if feas and (best is None or res.fun < best.fun):
            best = res
    if best is None:
        raise RuntimeError("SLSQP found no feasible point")
    D, L = split(best.x)
    return dict(value=-best.fun, psi=psi(best.x),
                deltas=[D[off[i]:off[i + 1]] for i in range(m)],
                lambdas=[L[off[i]:off[i + 1]] for i in range(m)])
