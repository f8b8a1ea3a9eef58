"""CPU fp64 oracle for the CFMM optimal-routing hot path  --  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product package
(``cfmm_routing_code_b200``) never does; its CUDA path fails loudly instead of
falling back to anything in here.

PARITY UNPINNED BY THE REFERENCE: the reference (angeris/cfmm-routing-code) has no
tests and no recorded outputs, and its only numerical engine (cvxpy + ECOS/Clarabel,
unpinned, ``README.md:6-9``) is not installable in this image.  What pins this oracle
instead: (1) ``oracle/primal_scipy.py`` solves the reference's *primal* program
(``arbitrage.py:50-82`` literally: variables Delta_i, Lambda_i >= 0, the three phi
constraints, the utility) with scipy SLSQP, an independent method; (2) zero duality
gap between that primal point and this module's dual point; (3) closed-form /
brute-force cross-checks in ``tests/test_oracle.py``.

What is restated here (reference file:line):
  * problem data layout: ``local_indices`` / ``reserves`` / ``fees``     arbitrage.py:5-28
  * net trade  psi = sum_i A_i (Lambda_i - Delta_i)                      arbitrage.py:42-54
  * post-trade reserves  R + gamma*Delta - Lambda (fee on Delta only)    arbitrage.py:60
  * phi = weighted geometric mean  (cp.geo_mean(x, p=w))                 arbitrage.py:65
  * phi = constant product        (cp.geo_mean(x) on 2 tokens)           arbitrage.py:68-70
  * phi = constant sum, plus new_reserves >= 0                           arbitrage.py:73-74
  * utilities: arbitrage  max c'psi, psi>=0                              arbitrage.py:57,77
               liquidation max psi[t], psi_j = -a_j                      liquidation.py:57,77-80
               swap        max psi[out], psi + t e_in >= 0               two-asset.py:66,86

The reference hands this program to an interior-point solver.  The oracle (and the
CUDA product) solve the same program by dual decomposition on the token price vector
nu: g(nu) = sum_j (nu_j - c_j) a_j + sum_i arb_i(A_i' nu), where arb_i is the optimal
arbitrage value of pool i at local prices, and grad g = a + psi(nu).
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional, Sequence

import numpy as np

KIND_GEOMEAN = 0   # weighted geometric mean; constant product is the (1/2, 1/2) 2-token case
KIND_CONST_SUM = 1
KIND_BOUNDED = 3   # constant product on virtual reserves R + o with real reserves >= 0 (one Uniswap-v3 tick range):
                   # not in the reference; the "more trading functions" extension point of arbitrage.py:63-74

_TINY = 1e-300
DT_MAX = 3.0      # largest log-price change of one Newton step
LM_SHIFTS = (1e-14, 1e-8, 1e-6, 1e-4, 1e-2, 1.0)     # damping ladder of the Newton system, times the mean diagonal


# --------------------------------------------------------------------------------------
# problem container (CSR over pools) -- replaces the dense A_i of arbitrage.py:42-48
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class Pools:
    n_tokens: int
    pool_ptr: np.ndarray   # int64 [m+1]
    tok_idx: np.ndarray    # int32 [nnz]   local_indices, concatenated
    reserves: np.ndarray   # f64   [nnz]
    weights: np.ndarray    # f64   [nnz]   per-slot parameter of the trading function: weight normalised per pool
                           #               (geomean), unused (const-sum), virtual-reserve offset o_j (bounded product)
    gamma: np.ndarray      # f64   [m]     fees[i] = 1 - fee
    kind: np.ndarray       # uint8 [m]

    @property
    def m(self) -> int:
        return len(self.gamma)

    def arity(self) -> np.ndarray:
        return np.diff(self.pool_ptr)

    @staticmethod
    def from_lists(n_tokens, local_indices, reserves, fees, kinds, weights=None) -> "Pools":
        """Build from the reference's literals (arbitrage.py:5-28).  ``kinds[i]`` is
        'geomean' | 'product' | 'sum'; ``weights[i]`` the cvxpy ``p=`` vector or None."""
        ptr = [0]
        idx, res, wts, kd = [], [], [], []
        for i, l in enumerate(local_indices):
            k = len(l)
            ptr.append(ptr[-1] + k)
            idx += list(l)
            res += [float(x) for x in reserves[i]]
            kk = kinds[i]
            if kk in ("sum", KIND_CONST_SUM):
                kd.append(KIND_CONST_SUM)
                wts += [0.0] * k
            elif kk in ("bounded_product", KIND_BOUNDED):
                kd.append(KIND_BOUNDED)
                wts += [float(x) for x in weights[i]]        # the virtual offsets, as given
            else:
                kd.append(KIND_GEOMEAN)
                w = np.ones(k) if (weights is None or weights[i] is None) else np.asarray(weights[i], float)
                wts += list(w / w.sum())     # cvxpy geo_mean normalises p to sum 1
        return Pools(int(n_tokens), np.asarray(ptr, np.int64), np.asarray(idx, np.int32),
                     np.asarray(res, np.float64), np.asarray(wts, np.float64),
                     np.asarray(fees, np.float64), np.asarray(kd, np.uint8))


@dataclasses.dataclass
class Utility:
    """U(psi) = c'psi - I{psi_j + a_j >= 0 (free_nu=False, eq=False), = 0 (eq), free (pinned)}.
    Dual: g(nu) = sum_j (nu_j - c_j) a_j + sum_i arb_i, over nu_j >= c_j | nu_j free | nu_j = c_j."""
    c: np.ndarray        # f64 [n]
    a: np.ndarray        # f64 [n]
    eq: np.ndarray       # bool [n]   psi_j + a_j == 0   (nu_j free, > 0)
    pinned: np.ndarray   # bool [n]   psi_j unconstrained (nu_j = c_j)

    @staticmethod
    def arbitrage(market_value) -> "Utility":           # arbitrage.py:57,77
        c = np.asarray(market_value, float)
        n = len(c)
        return Utility(c, np.zeros(n), np.zeros(n, bool), np.zeros(n, bool))

    @staticmethod
    def liquidate(n, target, basket) -> "Utility":      # liquidation.py:57,77-80
        c = np.zeros(n); c[target] = 1.0
        a = np.asarray(basket, float).copy(); a[target] = 0.0
        eq = np.ones(n, bool); eq[target] = False
        pinned = np.zeros(n, bool); pinned[target] = True
        return Utility(c, a, eq, pinned)

    @staticmethod
    def swap(n, tok_in, tok_out, t) -> "Utility":       # two-asset.py:41-45,66,86
        c = np.zeros(n); c[tok_out] = 1.0
        a = np.zeros(n); a[tok_in] = float(t)
        return Utility(c, a, np.zeros(n, bool), np.zeros(n, bool))


# --------------------------------------------------------------------------------------
# per-pool optimal arbitrage, scalar restatements (small cases; the definitional form)
# --------------------------------------------------------------------------------------
def arb_geomean_scalar(R, w, gamma, nu):
    """max nu'(L-D) s.t. prod (R+gamma D-L)^w >= prod R^w, D,L>=0   (arbitrage.py:60,65).
    KKT: x_j = clip(R_j, gamma*M*w_j/nu_j, M*w_j/nu_j); in logs h(s)=0 with s = log M,
    h(s) = sum_j w_j [max(s-tA_j,0) + min(s-tB_j,0)],  tB_j = log(R_j nu_j / w_j),
    tA_j = tB_j - log gamma.  Exact root by walking the sorted breakpoints."""
    R = np.asarray(R, float); w = np.asarray(w, float); nu = np.asarray(nu, float)
    w = w / w.sum()
    tB = np.log(R * nu / w)
    tA = tB - np.log(gamma)
    k = len(R)
    zero = np.zeros(k)
    if tB.max() <= tA.min():            # no-trade cone
        return zero, zero.copy(), -np.inf
    h = lambda s: float(np.sum(w * (np.maximum(s - tA, 0) + np.minimum(s - tB, 0))))
    bps = np.sort(np.concatenate([tA, tB]))
    s = None
    for p in range(len(bps) - 1):
        hl, hr = h(bps[p]), h(bps[p + 1])
        if hl <= 0.0 <= hr:
            s = bps[p] if hr == hl else bps[p] - hl * (bps[p + 1] - bps[p]) / (hr - hl)
            break
    assert s is not None
    D = R * np.expm1(np.maximum(s - tA, 0)) / gamma
    L = -R * np.expm1(np.minimum(s - tB, 0))
    return D, L, s


def arb_product_scalar(R, gamma, nu):
    """Closed form for sqrt(x1 x2) >= sqrt(R1 R2)  (arbitrage.py:68-70)."""
    D = np.zeros(2); L = np.zeros(2)
    p0, p1 = nu[0] * R[0], nu[1] * R[1]
    if gamma * p1 > p0:       # tender token 0, receive token 1
        t = np.sqrt(gamma * p1 / p0)
        D[0] = R[0] * (t - 1) / gamma; L[1] = R[1] * (1 - 1 / t)
    elif gamma * p0 > p1:
        t = np.sqrt(gamma * p0 / p1)
        D[1] = R[1] * (t - 1) / gamma; L[0] = R[0] * (1 - 1 / t)
    return D, L


def arb_bounded_product_scalar(R, o, gamma, nu):
    """sqrt((x0+o0)(x1+o1)) >= sqrt((R0+o0)(R1+o1)), x = R + gamma D - L >= 0: a constant-product curve on the virtual
    reserves V = R + o of which only the real part R can be paid out (a Uniswap-v3 position inside its tick range).
    Optimal trade = the constant-product one on V, with the payout capped at R_b; at the cap the tender follows from
    the curve: (V_a + gamma D_a)(V_b - R_b) = V_a V_b."""
    R = np.asarray(R, float); o = np.asarray(o, float)
    V = R + o
    D = np.zeros(2); L = np.zeros(2)
    p = nu * V
    for a, b in ((0, 1), (1, 0)):
        if gamma * p[b] > p[a]:
            t = np.sqrt(gamma * p[b] / p[a])
            L[b] = V[b] * (1 - 1 / t); D[a] = V[a] * (t - 1) / gamma
            if L[b] > R[b]:
                L[b] = R[b]; D[a] = V[a] * R[b] / (o[b] * gamma)
    return D, L


def _order(z, R, theta_bar, eps):
    """One constant-sum limit order (tender a, receive up to R of b), proximal-multiplier form.
    psi(z) = max_{0<=th<=R} th*z - (th-theta_bar)^2/(2 sigma), sigma = R/eps, z = gamma nu_b/nu_a - 1.
    Returns psi, th = psi'(z), psi''(z).  eps=0: the exact bang-bang LP (th = R if z>0 else 0)."""
    z = np.asarray(z, float)
    if eps <= 0:
        th = np.where(z > 0, R, 0.0)
        return th * z, th, np.zeros_like(z)
    sigma = R / eps
    th = np.clip(theta_bar + sigma * z, 0.0, R)
    psi = th * z - (th - theta_bar) ** 2 / (2.0 * sigma)
    curv = np.where((th > 0) & (th < R), sigma, 0.0)
    return psi, th, curv


def arb_sum_scalar(R, gamma, nu, eps=0.0, theta_bar=(0.0, 0.0)):
    """2-token constant-sum pool with x >= 0 (arbitrage.py:73-74): an LP, bang-bang in nu.
    Each direction is a limit order: tender a, receive up to R_b of b, gamma per unit.
    eps>0 is the proximal-multiplier smoothing (see _order); theta_bar[b] is the multiplier
    (= current fill estimate) of the order that pays out token b.  The smoothed arb value
    (nu_a/gamma) psi(r-1) is the perspective of a convex function: convex, degree 1, and its
    gradient (L - D) always satisfies the pool's own constraints (it only over-pays)."""
    D = np.zeros(2); L = np.zeros(2)
    for a, b in ((0, 1), (1, 0)):
        r = gamma * nu[b] / nu[a]
        psi, th, _ = _order(r - 1.0, float(R[b]), float(theta_bar[b]), eps)
        L[b] += float(th)
        D[a] += float((r * th - psi) / gamma)
    return D, L


# --------------------------------------------------------------------------------------
# bucketed, vectorised evaluation (same math; used for 1e4..1e6 pools and as cpu baseline)
# --------------------------------------------------------------------------------------
class Buckets:
    """Pools regrouped by (kind, arity) so each group is a dense (m_g, k) array."""

    def __init__(self, pools: Pools):
        self.pools = pools
        ar = pools.arity()
        self.groups = []
        keys = sorted(set(zip(pools.kind.tolist(), ar.tolist())))
        for kd, k in keys:
            sel = np.nonzero((pools.kind == kd) & (ar == k))[0]
            off = pools.pool_ptr[sel][:, None] + np.arange(k)[None, :]
            g = dict(kind=kd, k=k, sel=sel, off=off,
                     idx=pools.tok_idx[off].astype(np.int64), R=pools.reserves[off],
                     w=pools.weights[off], gamma=pools.gamma[sel])
            if kd == KIND_GEOMEAN:
                g["c"] = np.log(g["R"] / g["w"])
                g["lg"] = np.log(g["gamma"])
                g["is_cp"] = bool(k == 2 and np.all(g["w"] == 0.5))
            self.groups.append(g)


def _geomean_group(g, nu, lognu):
    idx, R, w, gam = g["idx"], g["R"], g["w"], g["gamma"]
    if g.get("is_cp"):
        n0, n1 = nu[idx[:, 0]], nu[idx[:, 1]]
        p0, p1 = n0 * R[:, 0], n1 * R[:, 1]
        f = gam * p1 > p0                      # 0 -> 1
        b = gam * p0 > p1                      # 1 -> 0
        q = np.where(f, gam * p1 / p0, np.where(b, gam * p0 / p1, 1.0))
        t = np.sqrt(q)
        D = np.zeros_like(R); L = np.zeros_like(R)
        D[:, 0] = np.where(f, R[:, 0] * (t - 1) / gam, 0.0)
        L[:, 1] = np.where(f, R[:, 1] * (1 - 1 / t), 0.0)
        D[:, 1] = np.where(b, R[:, 1] * (t - 1) / gam, 0.0)
        L[:, 0] = np.where(b, R[:, 0] * (1 - 1 / t), 0.0)
        # Hessian coefficient M = 2 sqrt(k nu0 nu1 / gamma) on trading pools
        M = np.where(f | b, 2.0 * np.sqrt(p0 * p1 / gam), 0.0)
        act = np.stack([f | b, f | b], 1)
        return D, L, M, act
    tB = g["c"] + lognu[idx]
    tA = tB - g["lg"][:, None]
    trade = tB.max(1) > tA.min(1)
    T = np.concatenate([tA, tB], 1)                                   # (m, 2k)
    hT = (w[:, None, :] * (np.maximum(T[:, :, None] - tA[:, None, :], 0)
                           + np.minimum(T[:, :, None] - tB[:, None, :], 0))).sum(2)
    Tm = np.where(hT <= 0, T, -np.inf)
    p = Tm.argmax(1)
    rows = np.arange(len(p))
    sL, hL = T[rows, p], hT[rows, p]
    W = (w * ((sL[:, None] >= tA) | (sL[:, None] < tB))).sum(1)
    s = np.where(hL < 0, sL - hL / np.maximum(W, _TINY), sL)
    zA = np.where(trade[:, None], np.maximum(s[:, None] - tA, 0), 0.0)
    zB = np.where(trade[:, None], np.minimum(s[:, None] - tB, 0), 0.0)
    D = R * np.expm1(zA) / gam[:, None]
    L = -R * np.expm1(zB)
    act = (zA > 0) | (zB < 0)
    M = np.where(trade, np.exp(s), 0.0)
    return D, L, M, act


def _bounded_group(g, nu):
    idx, R, o, gam = g["idx"], g["R"], g["w"], g["gamma"]
    V = R + o
    D = np.zeros_like(R); L = np.zeros_like(R)
    p0, p1 = nu[idx[:, 0]] * V[:, 0], nu[idx[:, 1]] * V[:, 1]
    hcoef = np.zeros(len(gam))
    for a, b, pa, pb in ((0, 1, p0, p1), (1, 0, p1, p0)):
        go = gam * pb > pa
        t = np.sqrt(np.where(go, gam * pb / pa, 1.0))
        Lb = V[:, b] * (1 - 1 / t)
        cap = go & (Lb > R[:, b])
        L[:, b] = np.where(go, np.where(cap, R[:, b], Lb), 0.0)
        with np.errstate(divide="ignore", invalid="ignore"):
            Dcap = V[:, a] * R[:, b] / (o[:, b] * gam)
        D[:, a] = np.where(go, np.where(cap, Dcap, V[:, a] * (t - 1) / gam), 0.0)
        hcoef += np.where(go & ~cap, 0.5 * np.sqrt(p0 * p1 / gam), 0.0)     # = M/4 of the constant-product pool
    return D, L, hcoef


def _sum_group(g, nu, eps):
    assert g["k"] == 2, "constant-sum pools are 2-token (arbitrage.py:11,19)"
    idx, R, gam = g["idx"], g["R"], g["gamma"]
    tb = g.setdefault("theta_bar", np.zeros_like(R))
    D = np.zeros_like(R); L = np.zeros_like(R)
    hcoef = np.zeros(len(gam))
    for a, b in ((0, 1), (1, 0)):
        na, nb = nu[idx[:, a]], nu[idx[:, b]]
        r = gam * nb / na
        psi, th, curv = _order(r - 1.0, R[:, b], tb[:, b], eps)
        L[:, b] = th
        D[:, a] = (r * th - psi) / gam
        hcoef += curv * nb * r
    return D, L, hcoef


def evaluate(bk: Buckets, nu, eps=0.0, want_trades=False, want_hess=False):
    """One dual evaluation: psi(nu) = sum_i A_i(L_i - D_i), arb(nu) = sum_i nu_i'(L_i - D_i).
    Returns dict(psi, arb[, delta, lam (CSR order)][, hess dense n x n])."""
    pools = bk.pools
    n = pools.n_tokens
    nu = np.asarray(nu, float)
    lognu = np.log(nu)
    psi = np.zeros(n)
    arb = 0.0
    out = {}
    if want_trades:
        delta = np.zeros_like(pools.reserves); lam = np.zeros_like(pools.reserves)
    if want_hess:
        Hs = np.zeros((n, n))   # scaled: true Hessian = diag(1/nu) Hs diag(1/nu)
    for g in bk.groups:
        if g["kind"] == KIND_GEOMEAN:
            D, L, M, act = _geomean_group(g, nu, lognu)
        elif g["kind"] == KIND_BOUNDED:
            D, L, hc = _bounded_group(g, nu)
        else:
            D, L, hc = _sum_group(g, nu, eps)
        y = L - D
        np.add.at(psi, g["idx"].ravel(), y.ravel())
        arb += float(np.sum(nu[g["idx"]] * y))
        if want_trades:
            delta[g["off"].ravel()] = D.ravel(); lam[g["off"].ravel()] = L.ravel()
        if want_hess:
            idx = g["idx"]
            if g["kind"] == KIND_GEOMEAN:
                wa = g["w"] * act
                Wa = np.maximum(wa.sum(1), _TINY)
                np.add.at(Hs, (idx.ravel(), idx.ravel()), (M[:, None] * wa).ravel())
                blk = -(M / Wa)[:, None, None] * wa[:, :, None] * wa[:, None, :]
                ii = np.repeat(idx[:, :, None], g["k"], 2); jj = np.repeat(idx[:, None, :], g["k"], 1)
                np.add.at(Hs, (ii.ravel(), jj.ravel()), blk.ravel())
            else:
                i0, i1 = idx[:, 0], idx[:, 1]
                np.add.at(Hs, (i0, i0), hc); np.add.at(Hs, (i1, i1), hc)
                np.add.at(Hs, (i0, i1), -hc); np.add.at(Hs, (i1, i0), -hc)
    out["psi"] = psi; out["arb"] = arb
    if want_trades:
        out["delta"] = delta; out["lam"] = lam
    if want_hess:
        out["hess_scaled"] = Hs
    return out


# --------------------------------------------------------------------------------------
# outer solver: projected Newton on the (smoothed) dual, dense linear algebra (small n)
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class Result:
    value: float                 # = prob.value   (arbitrage.py:84)
    psi: np.ndarray              # = psi.value    (liquidation.py:87)
    deltas: List[np.ndarray]     # = deltas[i].value   (two-asset.py:97)
    lambdas: List[np.ndarray]    # = lambdas[i].value
    nu: np.ndarray
    dual_value: float
    gap: float                   # (dual - primal)/max(|dual|, tiny)
    primal_infeas: float         # sum_j nu_j * violation_j / |dual value|
    iters: int
    evals: int
    status: str


def _bounds(util: Utility, nu_floor):
    lb = np.where(util.eq, nu_floor, np.maximum(util.c, nu_floor))
    return lb


def dual_value(util: Utility, nu, arb):
    return float(np.dot(nu - util.c, util.a) + arb)


def residuals(util: Utility, nu, psi, dual):
    """primal objective c'psi and the value-weighted constraint violation relative to the dual value."""
    s = psi + util.a                      # must be >=0 (ineq) / ==0 (eq) / free (pinned)
    viol = np.where(util.pinned, 0.0, np.where(util.eq, np.abs(s), np.maximum(-s, 0.0)))
    infeas = float(np.dot(nu, viol)) / max(abs(dual), 1e-300)
    primal = float(np.dot(util.c, psi))
    return primal, infeas


def solve(pools: Pools, util: Utility, nu0=None, tol=1e-9, eps=0.1, eps_min=1e-4, eps_shrink=0.5,
          max_outer=60, max_inner=100, verbose=False) -> Result:
    """Method of multipliers on the constant-sum fills (outer, ramp width eps shrinking geometrically)
    around a projected (active-set) Newton method in log-price coordinates on the smooth dual g_t(nu)
    (inner), dense n x n Hessian -- fine for the oracle's sizes (n <= ~1000).  With no constant-sum
    pool there is one outer pass.  Same algorithm as the product's solver.py, restated in numpy."""
    bk = Buckets(pools)
    n = pools.n_tokens
    sum_groups = [g for g in bk.groups if g["kind"] == KIND_CONST_SUM]
    for g in sum_groups:
        g["theta_bar"] = np.zeros_like(g["R"])
    scale = np.maximum(np.abs(util.c).max(), 1.0)
    lb = _bounds(util, 1e-12 * scale)
    fixed = util.pinned.copy()
    if nu0 is None:
        pos = util.c[util.c > 0]
        nu = np.where(util.c > 0, util.c, np.median(pos) if len(pos) else 1.0)
    else:
        nu = np.asarray(nu0, float).copy()
    nu = np.maximum(nu, lb)
    nu[fixed] = util.c[fixed]
    evals = 0; iters = 0
    status = "max_iter"
    eps_t = eps if sum_groups else 0.0

    def G(nu_, eps_=None, **kw):
        nonlocal evals
        evals += 1
        ev = evaluate(bk, nu_, eps_t if eps_ is None else eps_, **kw)
        ev["g"] = dual_value(util, nu_, ev["arb"])
        return ev

    a_inf = float(np.abs(util.a).max(initial=0.0))

    def kkt(nu_, ev_, err_prev):
        """max of the value-weighted residual sum_free |nu_j (a_j+psi_j)| / |g| and the per-token one
        max_free |a_j+psi_j| / max(|a|_inf, max over constrained tokens |psi_j|): the reference constrains psi token
        by token (liquidation.py:77-80, arbitrage.py:77), so a cheap token must not hide a large residual (nor may
        the unconstrained output of an objective-only token set the scale)."""
        grad = util.a + ev_["psi"]
        thr = min(1e-2, max(1e-3 * (err_prev if np.isfinite(err_prev) else 1e-2), 1e-14))     # active-set width (see solver.py)
        near = (nu_ <= lb * (1 + thr)) & ~util.eq
        free = ~(fixed | (near & (grad > 0)))
        pg = np.where(free, nu_ * grad, 0.0)
        den = max(abs(ev_["g"]), 1e-3 * np.dot(nu_, np.abs(grad)), 1e-300)
        feas = np.abs(np.where(free, grad, 0.0)).max(initial=0.0) / max(a_inf, np.abs(np.where(fixed, 0.0, ev_["psi"])).max(initial=0.0), 1e-300)
        return max(np.abs(pg).sum() / den, feas), grad, free, pg

    err = np.inf
    move = 1.0
    failed_before = False
    for outer in range(max_outer):
        ev = G(nu, want_hess=True)
        inner_status = "max_iter"
        inner_tol = max(tol, min(1e-3, 1e-2 * move)) if sum_groups else tol
        err, grad, free, pg = kkt(nu, ev, err)
        for _ in range(max_inner):
            iters += 1
            if verbose:
                print(f"outer={outer} it={iters} g={ev['g']:.15g} err={err:.3e} free={free.sum()}")
            if err <= inner_tol:
                inner_status = "optimal"
                break
            Hs = ev["hess_scaled"][np.ix_(free, free)]
            rhs = -pg[free]
            dbar = max(np.trace(Hs) / max(free.sum(), 1), 1e-300)
            # (near-)singular free-set systems (e.g. every pool tying some free prices to the rest is saturated) give an
            # enormous step along the null directions: damp the system (Levenberg-Marquardt, shift mu * mean diagonal)
            # until the step is a sane price change; the null directions then get a scaled gradient step
            rung = 0
            while True:
                # climb the ladder from `rung` until the step is a sane price change
                while True:
                    try:
                        dtf = np.linalg.solve(Hs + LM_SHIFTS[rung] * dbar * np.eye(len(rhs)), rhs)
                    except np.linalg.LinAlgError:
                        dtf = np.full(len(rhs), np.inf)
                    if (np.all(np.isfinite(dtf)) and np.abs(dtf).max(initial=0.0) <= DT_MAX) or rung == len(LM_SHIFTS) - 1:
                        break
                    rung += 1
                dt = np.zeros(n); dt[free] = dtf
                if not np.all(np.isfinite(dt)) or np.dot(pg, dt) >= 0:
                    dt = -pg / max(np.abs(pg).max(), 1e-300)
                big = np.abs(dt).max()
                if big > DT_MAX:        # still too long after the largest shift: keep the direction, bound the step
                    dt *= DT_MAX / big
                alpha = 1.0
                g0 = ev["g"]
                ok = False
                for _ls in range(50):
                    nu_t = np.maximum(nu * np.exp(np.clip(alpha * dt, -20, 20)), lb)
                    nu_t[fixed] = util.c[fixed]
                    ev_t = G(nu_t, want_hess=True)
                    lin = np.dot(grad, nu_t - nu)
                    if _ls == 0:
                        lin1 = lin                 # predicted decrease of the FULL step
                    if ev_t["g"] <= g0 + 1e-4 * lin:
                        ok = True
                        break
                    if abs(ev_t["g"] - g0) <= 1e-13 * abs(g0) or abs(lin1) <= 1e-9 * abs(g0):   # g cannot resolve this step
                        if kkt(nu_t, ev_t, err)[0] < 0.99 * err:
                            ok = True
                            break
                        if alpha < 1e-3:
                            break
                    alpha *= 0.5
                # a failed search along a barely damped direction (null-space dominated: long step, no predicted gain):
                # damp harder and try again before giving up
                if ok or rung == len(LM_SHIFTS) - 1:
                    break
                rung += 1
            if not ok:
                inner_status = "stalled"
                break
            nu, ev = nu_t, ev_t
            err, grad, free, pg = kkt(nu, ev, err)
        if not sum_groups:
            status = inner_status
            break
        fin = G(nu, want_trades=True)
        exact = G(nu, 0.0)
        gap_now = (exact["g"] - float(np.dot(util.c, fin["psi"]))) / max(abs(exact["g"]), 1e-300)
        if verbose:
            print(f"outer={outer} eps={eps_t:.1e} last move={move:.3e} gap={gap_now:.3e}")
        if inner_status == "optimal" and err <= tol and abs(gap_now) <= tol:
            status = "optimal"          # the only certified exit
            break
        status = inner_status if inner_status != "optimal" else "max_iter"
        if inner_status != "optimal" and failed_before and eps_t <= eps_min:
            break       # ramp at its narrowest and two failed passes: the residual sits at the fp64 floor (ratio / eps)
        failed_before = inner_status != "optimal"
        move = 0.0
        for g in sum_groups:
            th = fin["lam"][g["off"]]                     # Lambda_b IS the fill of the order paying b
            move = max(move, float(np.max(np.abs(th - g["theta_bar"]) / g["R"])))
            g["theta_bar"] = th.copy()
        eps_t = max(eps_min, eps_t * eps_shrink)
    fin = G(nu, want_trades=True)
    exact = G(nu, 0.0)
    dval = exact["g"]
    primal, infeas = residuals(util, nu, fin["psi"], dval)
    gap = (dval - primal) / max(abs(dval), 1e-300)
    ptr = pools.pool_ptr
    deltas = [fin["delta"][ptr[i]:ptr[i + 1]].copy() for i in range(pools.m)]
    lambdas = [fin["lam"][ptr[i]:ptr[i + 1]].copy() for i in range(pools.m)]
    return Result(primal, fin["psi"], deltas, lambdas, nu, dval, gap, infeas, iters, evals, status)
