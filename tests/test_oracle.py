"""Oracle vs golden vectors, closed forms and brute force (CPU only)."""
import numpy as np
import pytest
from scipy import optimize

from cfmm_routing_code_b200 import instances as I
from oracle import cfmm_oracle as O
import helpers as H


def _solve(d, util):
    return O.solve(H.oracle_pools(H.host_pools(d)), util, tol=1e-10)


def test_arbitrage_matches_golden(golden):
    d = I.arbitrage_instance()
    r = _solve(d, O.Utility.arbitrage(d["market_value"]))
    g = golden["arbitrage"]
    assert r.status == "optimal"
    assert abs(r.value - g["value"]) <= 1e-8 * abs(g["value"])           # scipy-primal golden
    assert abs(r.value - golden["survey_8c"]["arbitrage"]) <= 1e-9 * 21.5   # SURVEY 8c (zero-gap) value
    np.testing.assert_allclose(r.psi, g["psi"], atol=2e-6)
    assert abs(r.gap) <= 1e-8 and r.primal_infeas <= 1e-8
    for i in range(5):
        np.testing.assert_allclose(r.deltas[i], g["deltas"][i], atol=5e-5)
        np.testing.assert_allclose(r.lambdas[i], g["lambdas"][i], atol=5e-5)


def test_liquidation_matches_golden(golden):
    d = I.liquidation_instance()
    r = _solve(d, O.Utility.liquidate(5, d["target"], d["current_assets"]))
    g = golden["liquidation"]
    assert r.status == "optimal"
    assert abs(r.value - g["value"]) <= 1e-8 * abs(g["value"])
    assert abs(r.value - golden["survey_8c"]["liquidation"]) <= 1e-9 * 16
    np.testing.assert_allclose(r.psi, g["psi"], atol=2e-6)


def test_two_asset_sweep_matches_golden(golden):
    d = I.two_asset_instance()
    P = H.oracle_pools(H.host_pools(d))
    for j in range(0, 50, 3):
        r = O.solve(P, O.Utility.swap(3, 0, 2, d["amounts"][j]), tol=1e-10)
        g = golden["two_asset"][j]
        assert abs(r.value - g["value"]) <= 1e-7 * max(abs(g["value"]), 1), j
    assert abs(golden["two_asset"][0]["value"] - golden["survey_8c"]["two_asset_t0"]) < 1e-8
    assert abs(golden["two_asset"][49]["value"] - golden["survey_8c"]["two_asset_t50"]) < 1e-8


def test_product_closed_form_equals_geomean_breakpoint_solver():
    rng = np.random.default_rng(0)
    for _ in range(200):
        R = np.exp(rng.normal(3, 1, 2)); nu = np.exp(rng.normal(0, 1, 2)); gam = rng.choice([0.997, 0.999, 0.95])
        D1, L1 = O.arb_product_scalar(R, gam, nu)
        D2, L2, _ = O.arb_geomean_scalar(R, [0.5, 0.5], gam, nu)
        np.testing.assert_allclose(D1, D2, rtol=1e-9, atol=1e-12 * R.max())
        np.testing.assert_allclose(L1, L2, rtol=1e-9, atol=1e-12 * R.max())


def test_geomean_scalar_is_optimal_vs_brute_force():
    """The KKT solution beats random feasible trades and matches SLSQP on the pool's own program."""
    rng = np.random.default_rng(1)
    for k in (2, 3, 5):
        R = np.exp(rng.normal(2, 0.5, k)); w = rng.dirichlet(np.ones(k)); nu = np.exp(rng.normal(0, 0.3, k))
        gam = 0.99
        D, L, _ = O.arb_geomean_scalar(R, w, gam, nu)
        x = R + gam * D - L
        assert np.dot(w, np.log(x)) >= np.dot(w, np.log(R)) - 1e-12      # feasible (arbitrage.py:65)
        val = np.dot(nu, L - D)
        f = lambda z: -np.dot(nu, z[k:] - z[:k])
        con = dict(type="ineq", fun=lambda z: np.dot(w, np.log(np.maximum(R + gam * z[:k] - z[k:], 1e-300))
                                                     - np.log(R)))
        res = optimize.minimize(f, np.zeros(2 * k), method="SLSQP", bounds=[(0, None)] * (2 * k),
                                constraints=[con], options=dict(ftol=1e-14, maxiter=500))
        assert val >= -res.fun - 1e-7 * max(1, abs(val))


def test_no_trade_cone_and_homogeneity():
    R = np.array([10.0, 20.0]); gam = 0.997
    nu = np.array([2.0, 1.0])                   # pool price R1/R0 = 2 = nu0/nu1: inside the cone
    D, L = O.arb_product_scalar(R, gam, nu)
    assert not D.any() and not L.any()
    D, L, _ = O.arb_geomean_scalar(R, [0.5, 0.5], gam, nu)
    assert not D.any() and not L.any()
    nu = np.array([2.3, 1.0])
    D1, L1 = O.arb_product_scalar(R, gam, nu)
    D2, L2 = O.arb_product_scalar(R, gam, 7.5 * nu)
    np.testing.assert_allclose(D1, D2, rtol=1e-13); np.testing.assert_allclose(L1, L2, rtol=1e-13)


def test_const_sum_lp_rule_and_smoothing_is_pool_feasible():
    R = np.array([10.0, 10.0]); gam = 0.999
    D, L = O.arb_sum_scalar(R, gam, np.array([1.0, 1.01]))      # token 1 dearer: drain it
    assert L[1] == 10.0 and abs(D[0] - 10.0 / gam) < 1e-12 and L[0] == 0 and D[1] == 0
    D, L = O.arb_sum_scalar(R, gam, np.array([1.0, 1.0005]))    # inside the fee band: nothing
    assert not D.any() and not L.any()
    for thb in (0.0, 3.0):
        for r in np.linspace(0.999, 1.003, 9):
            nu = np.array([1.0, r / gam])
            D, L = O.arb_sum_scalar(R, gam, nu, eps=1e-3, theta_bar=(0.0, thb))
            x = R + gam * D - L
            assert x.sum() >= R.sum() - 1e-12 and np.all(x >= -1e-12)      # arbitrage.py:73-74


def test_vectorised_evaluation_equals_scalar_solvers():
    hp, s = H.mixed_host_pools(600, 40, seed=5)
    P = H.oracle_pools(hp)
    nu = H.random_prices(s["prices"], 7)
    ev = O.evaluate(O.Buckets(P), nu, eps=0.0, want_trades=True)
    ptr = P.pool_ptr
    for i in range(P.m):
        sl = slice(ptr[i], ptr[i + 1])
        R, w, loc = P.reserves[sl], P.weights[sl], nu[P.tok_idx[sl]]
        if P.kind[i] == O.KIND_CONST_SUM:
            D, L = O.arb_sum_scalar(R, P.gamma[i], loc)
        else:
            D, L, _ = O.arb_geomean_scalar(R, w, P.gamma[i], loc)
        np.testing.assert_allclose(ev["delta"][sl], D, rtol=1e-9, atol=1e-11 * R.max())
        np.testing.assert_allclose(ev["lam"][sl], L, rtol=1e-9, atol=1e-11 * R.max())


def test_gradient_and_hessian_by_finite_differences():
    hp, s = H.mixed_host_pools(300, 12, seed=3)
    hp.kind[:] = 0                      # geomean only: smooth
    hp.weights[hp.weights == 0] = 0.5
    P = H.oracle_pools(hp)
    bk = O.Buckets(P)
    nu = H.random_prices(s["prices"], 11, 0.2)
    ev = O.evaluate(bk, nu, want_hess=True)
    Hs = ev["hess_scaled"] / nu[:, None] / nu[None, :]
    for j in range(12):
        h = 1e-6 * nu[j]
        e = np.zeros(12); e[j] = h
        ep, em = O.evaluate(bk, nu + e), O.evaluate(bk, nu - e)
        assert abs((ep["arb"] - em["arb"]) / (2 * h) - ev["psi"][j]) <= 1e-5 * (abs(ev["psi"][j]) + 1)
        np.testing.assert_allclose((ep["psi"] - em["psi"]) / (2 * h), Hs[:, j],
                                   atol=2e-4 * np.abs(Hs[:, j]).max())


def test_c_restatement_matches_numpy_oracle():
    """oracle/cfmm_oracle_c.c (the multi-threaded CPU baseline of bench.py) against the numpy oracle"""
    from oracle import c_oracle as CO
    hp, s = H.cp_host_pools(50_000, 300, seed=2)
    nu = H.random_prices(s["prices"], 5)
    psi, arb, d, l = CO.eval_pairs(hp.tok_idx.reshape(-1, 2), hp.reserves.reshape(-1, 2), hp.gamma, 300, nu, True)
    ev = O.evaluate(O.Buckets(H.oracle_pools(hp)), nu, want_trades=True)
    np.testing.assert_allclose(psi, ev["psi"], atol=1e-12 * np.abs(ev["psi"]).max())
    assert abs(arb - ev["arb"]) <= 1e-12 * abs(ev["arb"])
    np.testing.assert_allclose(d.ravel(), ev["delta"], atol=1e-13 * hp.reserves.max())
    np.testing.assert_allclose(l.ravel(), ev["lam"], atol=1e-13 * hp.reserves.max())
    assert CO.num_threads() >= 1


def test_c_solve_matches_numpy_oracle_solve():
    """oracle_solve_pairs (the CPU arm's whole solve: Newton-PCG in C over the persistent pthread pool) vs O.solve"""
    from oracle import c_oracle as CO
    hp, s = H.cp_host_pools(10_000, 256, seed=0)
    idx, R = hp.tok_idx.reshape(-1, 2), hp.reserves.reshape(-1, 2)
    ro = O.solve(H.oracle_pools(hp), O.Utility.arbitrage(s["prices"]), tol=1e-10)
    for nt in (1, 3):                   # the pool is rebuilt when the thread count changes
        CO.set_threads(nt)
        nu, psi, res = CO.solve_pairs(idx, R, hp.gamma, 256, s["prices"], tol=1e-9)
        assert res.status == 0 and abs(res.primal_value - ro.value) <= 1e-8 * abs(ro.value)
        assert abs(res.gap) <= 1e-8 and res.primal_infeas <= 1e-8
        np.testing.assert_allclose(nu, ro.nu, rtol=1e-6)
    basket = I.synth_basket(256, s["prices"], seed=2)
    u = O.Utility.liquidate(256, 0, basket)
    nu0 = s["prices"] / s["prices"][0]
    ro = O.solve(H.oracle_pools(hp), u, nu0=nu0, tol=1e-10)
    nu, psi, res = CO.solve_pairs(idx, R, hp.gamma, 256, u.c, u.a, u.eq, u.pinned, nu0=nu0, tol=1e-9)
    assert res.status == 0 and abs(res.primal_value - ro.value) <= 1e-7 * abs(ro.value)
    np.testing.assert_allclose(psi[1:], -basket[1:], atol=1e-7 * basket.max())


# ---- property tests: random pools / prices, invariants of the per-pool solutions -------
from hypothesis import given, settings, strategies as st


@settings(max_examples=60, deadline=None)
@given(k=st.integers(2, 6), seed=st.integers(0, 10_000), gam=st.sampled_from([0.9, 0.99, 0.997, 0.9995]),
       spread=st.floats(0.0, 1.5))
def test_property_geomean_solution_satisfies_kkt(k, seed, gam, spread):
    """feasible on the trading-function boundary, never tenders and receives the same token, complementary
    slackness of the price bounds gamma*M*w/nu <= x <= M*w/nu, degree-0 in nu, and no better than no trade <=> zero"""
    rng = np.random.default_rng(seed)
    R = np.exp(rng.normal(3, 1, k)); w = rng.dirichlet(np.ones(k)); nu = np.exp(spread * rng.standard_normal(k))
    D, L, s = O.arb_geomean_scalar(R, w, gam, nu)
    assert np.all(D >= 0) and np.all(L >= 0) and np.all(D * L == 0)
    x = R + gam * D - L
    assert np.all(x > 0)
    assert abs(np.dot(w, np.log(x)) - np.dot(w, np.log(R))) <= 1e-11          # phi(x) == phi(R)  (arbitrage.py:65)
    val = float(np.dot(nu, L - D))
    assert val >= -1e-12 * np.dot(nu, R)                                         # at least as good as not trading
    if D.any() or L.any():
        M = np.exp(s)
        lo, hi = gam * M * w / nu, M * w / nu
        assert np.all(x >= lo * (1 - 1e-9)) and np.all(x <= hi * (1 + 1e-9))
        np.testing.assert_allclose(x[D > 0], lo[D > 0], rtol=1e-9)
        np.testing.assert_allclose(x[L > 0], hi[L > 0], rtol=1e-9)
    D2, L2, _ = O.arb_geomean_scalar(R, w, gam, 3.7 * nu)
    np.testing.assert_allclose(D2, D, rtol=1e-9, atol=1e-12 * R.max())
    np.testing.assert_allclose(L2, L, rtol=1e-9, atol=1e-12 * R.max())


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 10_000), gam=st.sampled_from([0.99, 0.997, 0.999]), eps=st.sampled_from([0.0, 1e-3, 1e-1]),
       frac=st.floats(0.0, 1.0))
def test_property_const_sum_trades_respect_the_pool(seed, gam, eps, frac):
    """any (smoothed or exact) constant-sum trade keeps sum(x) >= sum(R), x >= 0 (arbitrage.py:73-74) and D, L >= 0"""
    rng = np.random.default_rng(seed)
    R = np.exp(rng.normal(2, 1, 2)); nu = np.array([1.0, np.exp(0.01 * rng.standard_normal())])
    D, L = O.arb_sum_scalar(R, gam, nu, eps=eps, theta_bar=frac * R)
    x = R + gam * D - L
    assert np.all(D >= -1e-15) and np.all(L >= -1e-15)
    assert x.sum() >= R.sum() * (1 - 1e-13) and np.all(x >= -1e-12 * R.max())


def test_bounded_product_closed_form_vs_its_own_program():
    """constant product on virtual reserves with the payout capped by the real reserves (a v3 tick range): the closed
    form beats / matches SLSQP on the pool's own program, in range and at the cap; zero offsets = plain product"""
    rng = np.random.default_rng(4)
    for trial in range(12):
        R = np.exp(rng.normal(2, 0.5, 2)); o = np.exp(rng.normal(2, 1.0, 2)); gam = 0.997
        nu = np.exp(rng.normal(0, 0.6 if trial % 2 else 0.05, 2))
        D, L = O.arb_bounded_product_scalar(R, o, gam, nu)
        x = R + gam * D - L
        assert np.all(x >= -1e-12) and np.prod(x + o) >= np.prod(R + o) * (1 - 1e-12)
        f = lambda z: -np.dot(nu, z[2:] - z[:2])
        cons = [dict(type="ineq", fun=lambda z: np.sum(np.log(np.maximum(R + gam * z[:2] - z[2:] + o, 1e-300))
                                                      - np.log(R + o))),
                dict(type="ineq", fun=lambda z: R + gam * z[:2] - z[2:])]
        res = optimize.minimize(f, np.zeros(4), method="SLSQP", bounds=[(0, None)] * 4, constraints=cons,
                                options=dict(ftol=1e-14, maxiter=500))
        assert np.dot(nu, L - D) >= -res.fun - 1e-7 * max(1.0, abs(res.fun))
    D0, L0 = O.arb_bounded_product_scalar([10.0, 20.0], [0.0, 0.0], 0.997, np.array([2.3, 1.0]))
    D1, L1 = O.arb_product_scalar(np.array([10.0, 20.0]), 0.997, np.array([2.3, 1.0]))
    np.testing.assert_allclose(D0, D1, rtol=1e-14); np.testing.assert_allclose(L0, L1, rtol=1e-14)


def test_bounded_product_instance_dual_solution_matches_slsqp_primal():
    from cfmm_routing_code_b200 import instances as I
    from oracle import primal_scipy as PS
    d = I.v3_instance()
    P = O.Pools.from_lists(3, d["local_indices"], d["reserves"], d["fees"], d["kinds"], d["weights"])
    cases = [(O.Utility.arbitrage(d["market_value"]), d["market_value"], [("ge", j, 0.0) for j in range(3)]),
             (O.Utility.swap(3, 0, 2, 40.0), [0, 0, 1.0], [("ge", 0, 40.0), ("ge", 1, 0.0), ("ge", 2, 0.0)]),
             (O.Utility.swap(3, 1, 0, 25.0), [1.0, 0, 0], [("ge", 0, 0.0), ("ge", 1, 25.0), ("ge", 2, 0.0)])]
    for u, obj, cons in cases:
        r = O.solve(P, u, tol=1e-10)
        pr = PS.solve_primal(3, d["local_indices"], d["reserves"], d["fees"], d["kinds"], d["weights"], obj, cons)
        assert r.status == "optimal" and abs(r.gap) <= 1e-9
        assert abs(r.value - pr["value"]) <= 1e-7 * max(1.0, abs(r.value))
        np.testing.assert_allclose(r.psi, pr["psi"], atol=2e-5)
    # finite-difference check of gradient and scaled Hessian with bounded pools in and out of range
    bk = O.Buckets(P)
    nu = np.array([1.31, 1.02, 0.47])
    ev = O.evaluate(bk, nu, want_hess=True)
    Hs = ev["hess_scaled"] / nu[:, None] / nu[None, :]
    for j in range(3):
        h = 1e-6 * nu[j]; e = np.zeros(3); e[j] = h
        ep, em = O.evaluate(bk, nu + e), O.evaluate(bk, nu - e)
        assert abs((ep["arb"] - em["arb"]) / (2 * h) - ev["psi"][j]) <= 1e-5 * (abs(ev["psi"][j]) + 1)
        np.testing.assert_allclose((ep["psi"] - em["psi"]) / (2 * h), Hs[:, j], atol=2e-4 * np.abs(Hs).max())


def test_v3_position_helper_invariants():
    """instances.v3_position: (x + o_x)(y + o_y) = L^2, marginal price (y + o_y)/(x + o_x) = p, the real reserves are what
    is left when the price runs to either end of the range"""
    from cfmm_routing_code_b200 import instances as I
    for L, lo, hi, p in ((100.0, 0.8, 1.25, 1.0), (7.0, 1900.0, 2100.0, 2000.0), (3.0, 0.5, 0.6, 0.55)):
        (x, y), (ox, oy) = I.v3_position(L, lo, hi, p)
        assert x > 0 and y > 0
        assert abs((x + ox) * (y + oy) - L * L) <= 1e-12 * L * L
        assert abs((y + oy) / (x + ox) - p) <= 1e-12 * p
        assert abs((x + ox) - L / np.sqrt(p)) <= 1e-12 * L and abs(ox - L / np.sqrt(hi)) <= 1e-12 * L
        # draining token 1 completely moves the price to the lower end of the range: x_virtual = L / sqrt(lo)
        assert abs(L * L / oy - L / np.sqrt(lo)) <= 1e-12 * L / np.sqrt(lo)
    (x, y), _ = I.v3_position(5.0, 1.25, 1.6, 1.0)          # price below the range: all in token 0
    assert y == 0.0 and x > 0
    (x, y), _ = I.v3_position(5.0, 0.5, 0.8, 1.0)           # above: all in token 1
    assert x == 0.0 and y > 0
