"""CUDA path vs the CPU oracle, through the C ABI (libcfmm_b200.so).  Run on the B200 box: pytest -m gpu."""
import ctypes as C

import numpy as np
import pytest
import torch

import cfmm_routing_code_b200 as cf
from cfmm_routing_code_b200 import _lib, instances as I
from oracle import cfmm_oracle as O
import helpers as H

pytestmark = pytest.mark.gpu

F64 = dict(dtype=torch.float64, device="cuda")


def _oracle_eval(hp, nu, eps=0.0, theta=None, **kw):
    bk = O.Buckets(H.oracle_pools(hp))
    if theta is not None:
        for g in bk.groups:
            if g["kind"] == O.KIND_CONST_SUM:
                g["theta_bar"] = theta[g["off"]]
    return O.evaluate(bk, nu, eps, **kw)


def _check_eval(hp, nu, eps=0.0, theta=None, scatter_mode=0, layout="blocked"):
    st = cf.PoolStore(hp, layout=layout)
    st.lib.cfmm_set_scatter_mode(scatter_mode)
    try:
        if theta is not None:
            for b in st.buckets:
                if b.kind == _lib.KIND_SUM:
                    b.theta_bar[:, :b.m].copy_(torch.as_tensor(theta[b.off], **F64))
        acc = st.evaluate(torch.as_tensor(nu, **F64), eps, trades=True, hess=True).cpu().numpy()
        d, l = st.gather_trades()
    finally:
        st.lib.cfmm_set_scatter_mode(0)
    ref = _oracle_eval(hp, nu, eps, theta, want_trades=True, want_hess=True)
    Rmax = np.maximum.reduceat(hp.reserves, hp.pool_ptr[:-1])
    Rrep = np.repeat(Rmax, np.diff(hp.pool_ptr))
    # bit-level agreement is not expected (sqrt/exp/sum order); 1e-12 of the reserve scale is
    assert np.max(np.abs(d - ref["delta"]) / Rrep) <= 1e-12
    assert np.max(np.abs(l - ref["lam"]) / Rrep) <= 1e-12
    gross = np.zeros(hp.n_tokens)
    np.add.at(gross, hp.tok_idx, ref["delta"] + ref["lam"])
    assert np.max(np.abs(acc[:-1] - ref["psi"]) / (gross + 1e-300 + 1e-9 * gross.max())) <= 1e-11
    assert abs(acc[-1] - ref["arb"]) <= 1e-11 * np.dot(nu, gross)
    return st, ref


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_product_eval_matches_oracle(mode):
    """plain layout: LDG kernel with global / shared-memory scatter, and the TMA-staged kernel"""
    hp, s = H.cp_host_pools(20_000, 97, seed=11)
    _check_eval(hp, H.random_prices(s["prices"], 1), scatter_mode=mode, layout="plain")


@pytest.mark.parametrize("m,n", [(20_000, 97), (1500, 700), (896, 8), (897, 3000), (70_000, 4096)])
def test_blocked_product_eval_matches_oracle(m, n):
    """token-blocked layout (csrc/cfmm_blocked.cu): ragged last tile, few and many tokens per tile"""
    hp, s = H.cp_host_pools(m, n, seed=13)
    st, _ = _check_eval(hp, H.random_prices(s["prices"], 1))
    assert any(getattr(b, "blocked", False) for b in st.buckets)


def test_blocked_hvp_and_diag_match_oracle_over_several_tiles_per_cta():
    for m, n in ((20_000, 97), (70_000, 4096), (400_000, 512)):       # the last: several tiles per CTA
        hp, s = H.cp_host_pools(m, n, seed=m % 97)
        st, ref = _check_eval(hp, H.random_prices(s["prices"], 1))
        assert st.buckets[0].c_blocked.pools_per_tile == 896
        v = np.random.default_rng(2).standard_normal(n)
        Hs = ref["hess_scaled"]
        np.testing.assert_allclose(st.hvp(torch.as_tensor(v, **F64)).cpu().numpy(), Hs @ v,
                                   atol=1e-10 * np.abs(Hs).max())
        np.testing.assert_allclose(st.hess_diag().cpu().numpy(), np.diag(Hs), atol=1e-11 * np.abs(Hs).max())


def test_native_layout_builder_tables_satisfy_the_layout_invariants():
    """cfmm_blocked_build (csrc/cfmm_layout.cu: keys + radix sort + one CTA per tile): the same invariants the CPU test
    checks for the torch builder -- emulated scatter == index_add, rows longest first, ids map back to tokens -- plus
    agreement of the slabs with the pools' own data, for ragged last tiles, tiny and many-token problems"""
    from cfmm_routing_code_b200 import pools as PL
    lib = _lib.load()
    P, rs, ts, cap = PL.blocked_layout_info(lib)
    for m, n in ((5000, 300), (700, 3), (40_000, 2000), (896, 8), (897, 300), (200_000, 4096)):
        hp, s = H.cp_host_pools(m, n, seed=m % 13)
        st = cf.PoolStore(hp)
        b = st.buckets[0]
        assert b.blocked and b.tables["tok_per_tile"] is None and not len(b.residual)       # the native path built it
        order = b.order.cpu().to(torch.int64)
        assert sorted(order.tolist()) == list(range(m))                                     # a permutation of the pools
        t = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in b.tables.items()}
        idx = torch.as_tensor(hp.tok_idx.reshape(-1, 2).T.astype(np.int64).copy())
        H.check_blocked_tables(t, idx, order, n, P, rs, ts, cap)
        R = torch.as_tensor(hp.reserves.reshape(-1, 2))
        assert torch.equal(b.r0[:m].cpu(), R[order, 0]) and torch.equal(b.r1[:m].cpu(), R[order, 1])
        np.testing.assert_allclose(b.gamma_inv[:m].cpu().numpy(), 1.0 / hp.gamma[order.numpy()], rtol=1e-15)
        assert bool((b.r0[m:] == 1).all() and (b.gamma_inv[m:] == 1).all())                 # padding pools: inert
    with pytest.raises(ValueError):                                                          # validation lives in the key kernel
        cf.PoolStore(cf.HostPools.from_pairs(2, [[0, 0]] * 4, [[1.0, 1.0]] * 4, [0.99] * 4))
    # a tile that would touch more tokens than a tile may (897 pools over 3000 tokens): the native builder reports it and
    # the general builder takes over (blocked part + plain residual bucket); results are checked against the oracle in
    # test_blocked_product_eval_matches_oracle[897-3000]
    hp, s = H.cp_host_pools(897, 3000, seed=1)
    st = cf.PoolStore(hp)
    assert sum(b.m for b in st.buckets) == 897 and any(not getattr(b, "blocked", False) for b in st.buckets)


def test_blocked_layout_falls_back_when_tiles_touch_too_many_tokens():
    """every pool on its own pair of tokens: no tile can stay under the per-tile token cap -> plain bucket"""
    m = 4 * 896                         # whole tiles only (a ragged last tile of few pools would stay blocked)
    idx = np.arange(2 * m).reshape(m, 2)
    rng = np.random.default_rng(3)
    hp = cf.HostPools.from_pairs(2 * m, idx, np.exp(rng.normal(3, 1, (m, 2))), np.full(m, 0.997))
    st, _ = _check_eval(hp, np.exp(rng.normal(0, 0.5, 2 * m)))
    assert not any(getattr(b, "blocked", False) for b in st.buckets)
    # mixed: a dense core plus the sparse fringe
    hp2, s2 = H.cp_host_pools(8000, 64, seed=5)
    idx2 = np.concatenate([hp2.tok_idx.reshape(-1, 2), 64 + idx])
    R2 = np.concatenate([hp2.reserves.reshape(-1, 2), hp.reserves.reshape(-1, 2)])
    hp3 = cf.HostPools.from_pairs(64 + 2 * m, idx2, R2, np.concatenate([hp2.gamma, hp.gamma]))
    _check_eval(hp3, np.exp(rng.normal(0, 0.3, 64 + 2 * m)))


def test_blocked_hvp_and_diag_match_plain():
    hp, s = H.cp_host_pools(30_000, 300, seed=17)
    nu = torch.as_tensor(H.random_prices(s["prices"], 2), **F64)
    v = torch.randn(300, **F64)
    outs = []
    for layout in ("blocked", "plain"):
        st = cf.PoolStore(hp, layout=layout)
        st.evaluate(nu, hess=True)
        outs.append((st.hvp(v).clone(), st.hess_diag().clone(), st.hess_dense().clone()))
    for x, y in zip(*outs):
        assert float((x - y).abs().max()) <= 1e-11 * float(y.abs().max())


def test_product_eval_extreme_prices_and_no_trade():
    hp, s = H.cp_host_pools(4096, 31, seed=12)
    _check_eval(hp, s["prices"] * np.exp(3.0 * np.random.default_rng(0).standard_normal(31)))
    # exactly the pool-implied prices: inside every no-trade cone -> psi == 0
    hp2 = cf.HostPools.from_pairs(2, np.array([[0, 1]] * 8), np.array([[10.0, 20.0]] * 8), np.full(8, 0.997))
    st = cf.PoolStore(hp2)
    acc = st.evaluate(torch.tensor([2.0, 1.0], **F64)).cpu().numpy()
    assert not acc.any()


@pytest.mark.parametrize("eps", [0.0, 1e-3])
def test_mixed_eval_matches_oracle(eps):
    hp, s = H.mixed_host_pools(6000, 120, seed=21)
    rng = np.random.default_rng(2)
    theta = rng.random(len(hp.reserves)) * hp.reserves * (eps > 0)
    _check_eval(hp, H.random_prices(s["prices"], 3, 0.03), eps=eps, theta=theta)


def test_bounded_product_bucket_matches_oracle():
    """the bounded-liquidity product (v3 tick range) in the pool-parallel path: trades, psi, arb and the scaled
    Hessian of one evaluation, in range / at the payout cap / out of range, alone and next to every other kind"""
    d = I.v3_instance(); hp = H.host_pools(d)
    for nu in ([1.31, 1.02, 0.47], [1.0, 1.0, 0.5], [3.0, 1.0, 0.2], [0.3, 1.0, 2.0]):
        st, ref = _check_eval(hp, np.array(nu))
        Hs = ref["hess_scaled"]
        np.testing.assert_allclose(st.hess_dense().cpu().numpy(), Hs, atol=1e-11 * max(np.abs(Hs).max(), 1e-300))
    rng = np.random.default_rng(17)
    for _ in range(6):
        hp, d, prices = H.random_small_problem(rng)
        st, ref = _check_eval(hp, prices * np.exp(0.1 * rng.standard_normal(hp.n_tokens)))
        Hs = ref["hess_scaled"]
        np.testing.assert_allclose(st.hess_dense().cpu().numpy(), Hs, atol=1e-11 * max(np.abs(Hs).max(), 1e-300))
        v = rng.standard_normal(hp.n_tokens)
        np.testing.assert_allclose(st.hvp(torch.as_tensor(v, **F64)).cpu().numpy(), Hs @ v,
                                   atol=1e-10 * max(np.abs(Hs).max(), 1e-300))


def test_random_small_problems_of_every_kind_through_the_pool_parallel_path():
    rng = np.random.default_rng(23)
    for _ in range(5):
        hp, d, prices = H.random_small_problem(rng)
        op = H.oracle_pools(hp)
        for u in H.random_utilities(rng, hp.n_tokens, prices):
            class _U:
                def spec(self, n, u=u):
                    return cf.DualSpec(u.c, u.a, u.eq, u.pinned)
            r = cf.solve_pools(hp, _U(), method="pools", tol=1e-8)
            ro = O.solve(op, u, tol=1e-8)
            assert abs(r.value - ro.value) <= 1e-7 * max(abs(ro.dual_value), 1e-300)
            assert abs(r.gap) <= 1e-7 and r.primal_infeas <= 1e-7


def test_large_arity_generic_kernel():
    rng = np.random.default_rng(5)
    n, m, k = 64, 300, 13
    idx = np.stack([rng.choice(n, k, replace=False) for _ in range(m)])
    hp = cf.HostPools.from_lists(n, idx.tolist(), np.exp(rng.normal(3, 1, (m, k))).tolist(),
                                 rng.choice([0.997, 0.99], m).tolist(), ["geomean"] * m,
                                 rng.dirichlet(np.ones(k), m).tolist())
    _check_eval(hp, np.exp(rng.normal(0, 0.2, n)))


def test_hessian_products_match_oracle():
    hp, s = H.mixed_host_pools(5000, 60, seed=31)
    nu = H.random_prices(s["prices"], 4, 0.03)
    rng = np.random.default_rng(6)
    theta = 0.5 * hp.reserves
    st, ref = _check_eval(hp, nu, eps=1e-2, theta=theta)
    Hs = ref["hess_scaled"]
    scale = np.abs(Hs).max()
    np.testing.assert_allclose(st.hess_dense().cpu().numpy(), Hs, atol=1e-11 * scale)
    np.testing.assert_allclose(st.hess_diag().cpu().numpy(), np.diag(Hs), atol=1e-11 * scale)
    v = rng.standard_normal(60)
    np.testing.assert_allclose(st.hvp(torch.as_tensor(v, **F64)).cpu().numpy(), Hs @ v, atol=1e-10 * scale)


@pytest.mark.parametrize("method", ["pools", "thread", "auto"])
def test_reference_instances_end_to_end(golden, method):
    """the three scripts' instances through the pool-parallel kernels under the outer loop ('pools') and through the
    one-thread-per-problem solver ('thread', what 'auto' picks at this size)"""
    d = I.arbitrage_instance()
    r = cf.solve(d["local_indices"], d["reserves"], d["fees"], d["kinds"], d["weights"],
                 utility=cf.Arbitrage(d["market_value"]), tol=1e-9, method=method)
    g = golden["arbitrage"]
    assert r.status == "optimal"
    assert abs(r.value - g["value"]) <= 1e-6 * abs(g["value"])          # BASELINE.md pass criterion
    assert abs(r.value - golden["survey_8c"]["arbitrage"]) <= 1e-8 * 21.5
    np.testing.assert_allclose(r.psi, g["psi"], atol=1e-6 * 3.25)
    for i in range(5):
        np.testing.assert_allclose(r.deltas[i], g["deltas"][i], atol=5e-5)
        np.testing.assert_allclose(r.lambdas[i], g["lambdas"][i], atol=5e-5)
    d = I.liquidation_instance()
    r = cf.solve(d["local_indices"], d["reserves"], d["fees"], d["kinds"], d["weights"],
                 utility=cf.Liquidate(d["target"], d["current_assets"]), tol=1e-9, method=method)
    g = golden["liquidation"]
    assert r.status == "optimal"
    assert abs(r.psi[4] - g["value"]) <= 1e-6 * g["value"]
    np.testing.assert_allclose(r.psi, g["psi"], atol=1e-6 * 15.9)
    d = I.two_asset_instance()
    for j in (0, 7, 15, 23, 31, 49):
        r = cf.solve(d["local_indices"], d["reserves"], d["fees"], d["kinds"], d["weights"],
                     utility=cf.Swap(0, 2, d["amounts"][j]), tol=1e-9, method=method)
        assert abs(r.value - golden["two_asset"][j]["value"]) <= 1e-6 * max(1.0, golden["two_asset"][j]["value"]), j


@pytest.mark.parametrize("batched", [True, False])
def test_two_asset_sweep_all_50_points_match_golden(golden, batched):
    """two-asset.py:40-100: u(t) and the per-pool flows (lambdas - deltas, :93-94) for every t in linspace(0, 50);
    batched = all 50 solves in one kernel launch, else one warm-started solve after the other"""
    d = I.two_asset_instance()
    rs = cf.solve_sweep(d["local_indices"], d["reserves"], d["fees"], d["kinds"], d["weights"],
                        [cf.Swap(d["tok_in"], d["tok_out"], t) for t in d["amounts"]], tol=1e-9, batched=batched)
    assert len(rs) == 50
    for j, r in enumerate(rs):
        g = golden["two_asset"][j]
        assert r.status == "optimal", j
        assert abs(r.value - g["value"]) <= 1e-6 * max(1.0, g["value"]), j
        np.testing.assert_allclose(r.psi, g["psi"], atol=2e-5)
    # u(t) is concave and increasing in the tendered amount
    u = np.array([r.value for r in rs])
    assert np.all(np.diff(u) > 0) and np.all(np.diff(u, 2) < 1e-6)


def test_back_to_back_evaluations_keep_their_buffers_consistent():
    """ping-pong psi buffers cleared inside the kernels + programmatic dependent launch: 60 alternating calls"""
    hp, s = H.cp_host_pools(60_000, 500, seed=23)
    st = cf.PoolStore(hp)
    nus = [torch.as_tensor(H.random_prices(s["prices"], k, 0.02), **F64) for k in range(3)]
    ref = [st.evaluate(nu).clone() for nu in nus]
    torch.cuda.synchronize()
    for it in range(60):
        k = (it * 7) % 3
        out = st.evaluate(nus[k], hess=(it % 5 == 0)).clone()
        assert float((out - ref[k]).abs().max()) <= 1e-9 * float(ref[k].abs().max()), it
    st.evaluate(nus[0], hess=True)
    v = torch.randn(500, **F64)
    y0 = st.hvp(v).clone()
    for _ in range(9):
        assert float((st.hvp(v) - y0).abs().max()) <= 1e-9 * float(y0.abs().max())


@pytest.mark.parametrize("linear_solver", ["cg", "dense"])
def test_cfg2_solve_matches_oracle(linear_solver):
    """BASELINE.json configs[1]: 10k constant-product pools, 256 tokens."""
    hp, s = H.cp_host_pools(10_000, 256, seed=0)
    r = cf.solve_pools(hp, cf.Arbitrage(s["prices"]), tol=1e-9, linear_solver=linear_solver)
    ro = O.solve(H.oracle_pools(hp), O.Utility.arbitrage(s["prices"]), tol=1e-10)
    assert r.status == "optimal" and ro.status == "optimal"
    assert abs(r.value - ro.value) <= 1e-8 * abs(ro.value)
    assert abs(r.gap) <= 1e-8 and r.primal_infeas <= 1e-8
    gross = np.zeros(256); np.add.at(gross, hp.tok_idx, np.concatenate(ro.deltas) + np.concatenate(ro.lambdas))
    assert np.max(np.abs(r.psi - ro.psi) / gross.max()) <= 1e-7


@pytest.mark.parametrize("impl", ["persist", "hostloop"])
def test_native_solver_matches_python_solver_and_oracle(impl):
    """the persistent solver kernel (cfmm_persist_solve: one launch per solve) and the C++ host loop (cfmm_blocked_solve)
    vs solver.py on the same store, cfg2 size, arbitrage and liquidation"""
    hp, s = H.cp_host_pools(10_000, 256, seed=0)
    st = cf.PoolStore(hp)
    r_nat = cf.solve_pools(hp, cf.Arbitrage(s["prices"]), tol=1e-9, store=st, native=impl)
    r_py = cf.solve_pools(hp, cf.Arbitrage(s["prices"]), tol=1e-9, store=st, native=False)
    ro = O.solve(H.oracle_pools(hp), O.Utility.arbitrage(s["prices"]), tol=1e-10)
    assert r_nat.status == "optimal" and r_py.status == "optimal"
    assert r_nat.info.history == [] and len(r_py.info.history) > 0          # really two different loops
    for r in (r_nat, r_py):
        assert abs(r.value - ro.value) <= 1e-8 * abs(ro.value)
        assert abs(r.gap) <= 1e-8 and r.primal_infeas <= 1e-8
    np.testing.assert_allclose(r_nat.nu, r_py.nu, rtol=1e-7)
    d_ref = np.concatenate(ro.deltas)
    np.testing.assert_allclose(np.concatenate(r_nat.deltas), d_ref, atol=1e-6 * np.abs(d_ref).max())
    basket = I.synth_basket(256, s["prices"], seed=2)
    nu0 = s["prices"] / s["prices"][0]
    r_nat = cf.solve_pools(hp, cf.Liquidate(0, basket), nu0=nu0, tol=1e-9, store=st, native=impl)
    ro = O.solve(H.oracle_pools(hp), O.Utility.liquidate(256, 0, basket), nu0=nu0, tol=1e-10)
    assert r_nat.status == "optimal" and abs(r_nat.value - ro.value) <= 1e-7 * abs(ro.value)
    np.testing.assert_allclose(r_nat.psi[1:], -basket[1:], atol=1e-7 * basket.max())


def test_persistent_and_hostloop_solvers_take_the_same_path():
    """same method, same constants: iteration / evaluation / HVP counts of the two native loops agree on a problem with
    several tiles per CTA (400k pools) and on tiny ones (a single tile; fewer tiles than CTAs), from several starts"""
    for m, n, seed in ((400_000, 2048, 5), (700, 12, 6), (30_000, 700, 7)):
        hp, s = H.cp_host_pools(m, n, seed=seed)
        st = cf.PoolStore(hp)
        for util, nu0 in ((cf.Arbitrage(s["prices"]), None),
                          (cf.Swap(0, 1, float(hp.reserves[0]) * 0.5), s["prices"] / s["prices"][1])):
            ra = cf.solve_pools(hp, util, nu0=nu0, tol=1e-8, store=st, native="persist", want_trades=False)
            rb = cf.solve_pools(hp, util, nu0=nu0, tol=1e-8, store=st, native="hostloop", want_trades=False)
            assert ra.status == rb.status == "optimal", (m, ra.status, rb.status)
            assert abs(ra.value - rb.value) <= 1e-9 * max(abs(rb.value), 1e-300) + 1e-12 * abs(rb.dual_value)
            # same method; summation orders differ (atomics, fused PCG recurrences), so the paths may part by an iteration or
            # two.  The host loop re-evaluates the current point after every rejected trial and, once g no longer resolves the
            # steps, can spend dozens of evaluations backtracking where the persistent kernel (KKT data of both points kept)
            # needs none: only an upper bound on the persistent solver's evaluations is asserted
            assert abs(ra.iters - rb.iters) <= 4 and ra.evals <= rb.evals + 6, (ra.iters, rb.iters, ra.evals, rb.evals)
            np.testing.assert_allclose(ra.nu, rb.nu, rtol=1e-6)


def test_cfg3_small_mixed_solve_matches_oracle():
    hp, s = H.mixed_host_pools(8000, 150, seed=1)
    r = cf.solve_pools(hp, cf.Arbitrage(s["prices"]), tol=1e-8)
    ro = O.solve(H.oracle_pools(hp), O.Utility.arbitrage(s["prices"]), tol=1e-9)
    assert r.status == "optimal"
    assert abs(r.value - ro.value) <= 1e-6 * abs(ro.value)
    assert abs(r.gap) <= 1e-6 and r.primal_infeas <= 1e-6


def test_cfg4_small_liquidation_matches_oracle():
    hp, s = H.mixed_host_pools(8000, 150, seed=2)
    basket = I.synth_basket(150, s["prices"], seed=2)
    nu0 = s["prices"] / s["prices"][0]
    r = cf.solve_pools(hp, cf.Liquidate(0, basket), nu0=nu0, tol=1e-8)
    ro = O.solve(H.oracle_pools(hp), O.Utility.liquidate(150, 0, basket), nu0=nu0, tol=1e-9)
    assert r.status == "optimal"
    assert abs(r.value - ro.value) <= 1e-6 * abs(ro.value)
    np.testing.assert_allclose(r.psi[1:], -basket[1:], atol=1e-6 * basket.max())


def test_full_size_mixed_configs_reach_a_certified_optimum():
    """BASELINE.json configs[2] and [3] at full size (100k mixed pools, 1000 tokens; the liquidation.py objective over
    them): oracle-free certificates -- weak duality, relative gap, value-weighted infeasibility AND every token's own
    residual (liquidation.py:77-80 constrains psi token by token) -- plus shard additivity of the mixed evaluation"""
    s = I.synth_mixed(100_000, 1000, seed=1)
    hp = cf.HostPools(1000, s["pool_ptr"], s["tok_idx"], s["reserves"], s["weights"], s["gamma"], s["kind"])
    st = cf.PoolStore(hp)
    r = cf.solve_pools(hp, cf.Arbitrage(s["prices"]), tol=1e-6, store=st, want_trades=False)
    assert r.status == "optimal" and abs(r.gap) <= 1e-6 and r.primal_infeas <= 1e-6
    assert r.dual_value >= r.value - 1e-6 * abs(r.dual_value)                       # weak duality
    assert r.psi.min() >= -1e-6 * np.abs(r.psi).max()                               # arbitrage.py:77, token by token
    assert np.all(r.nu >= s["prices"] * (1 - 1e-12))                                # dual box nu >= c
    nu = torch.as_tensor(H.random_prices(s["prices"], 5, 0.01), **F64)
    # (exact evaluation, eps = 0: the smoothed one depends on the fill multipliers, which `st` carries from the solve above)
    full = st.evaluate(nu, 0.0).clone()
    halves = sum(cf.PoolStore(hp, rank=k, world=2).evaluate(nu, 0.0).clone() for k in range(2))
    assert float((full - halves)[:-1].abs().max()) <= 1e-9 * float(full[:-1].abs().max())
    s = I.synth_mixed(100_000, 1000, seed=2)
    hp = cf.HostPools(1000, s["pool_ptr"], s["tok_idx"], s["reserves"], s["weights"], s["gamma"], s["kind"])
    basket = I.synth_basket(1000, s["prices"], seed=2)
    r = cf.solve_pools(hp, cf.Liquidate(0, basket), nu0=s["prices"] / s["prices"][0], tol=1e-6, want_trades=False)
    assert r.status == "optimal" and abs(r.gap) <= 1e-6 and r.primal_infeas <= 1e-6
    np.testing.assert_allclose(r.psi[1:], -basket[1:], atol=1e-6 * basket.max())    # psi_j + a_j == 0 for every basket token
    assert r.value > 0 and abs(r.value - r.psi[0]) <= 1e-12 * abs(r.value)


def test_full_size_properties_1m_pools():
    """BASELINE.json configs[4] size (1M pools, 4096 tokens): oracle-free invariants."""
    hp, s = H.cp_host_pools(1_000_000, 4096, seed=3)
    st = cf.PoolStore(hp)
    nu = torch.as_tensor(H.random_prices(s["prices"], 9, 0.01), **F64)
    a1 = st.evaluate(nu).clone()
    a2 = st.evaluate(7.0 * nu).clone()
    gross = float(a1[:-1].abs().max())
    assert float((a1[:-1] - a2[:-1]).abs().max()) <= 1e-9 * gross          # psi is degree 0
    assert abs(float(a2[-1]) - 7.0 * float(a1[-1])) <= 1e-10 * abs(float(a2[-1]))   # arb is degree 1
    assert abs(float(torch.dot(nu, a1[:-1])) - float(a1[-1])) <= 1e-9 * abs(float(a1[-1]))  # arb = nu'psi
    # shard additivity: two half-stores sum to the whole (what the multi-GPU all-reduce relies on)
    h0 = cf.PoolStore(hp, rank=0, world=2).evaluate(nu).clone()
    h1 = cf.PoolStore(hp, rank=1, world=2).evaluate(nu).clone()
    assert float((h0 + h1 - a1)[:-1].abs().max()) <= 1e-9 * gross
    # a 100k-pool slice against the oracle (the oracle finishes this in well under a second)
    sub = cf.HostPools.from_pairs(4096, hp.tok_idx.reshape(-1, 2)[:100_000], hp.reserves.reshape(-1, 2)[:100_000],
                                  hp.gamma[:100_000])
    _check_eval(sub, nu.cpu().numpy())


def test_full_size_solve_reaches_1e6_gap():
    hp, s = H.cp_host_pools(1_000_000, 4096, seed=3)
    r = cf.solve_pools(hp, cf.Arbitrage(s["prices"]), tol=1e-6, want_trades=False)
    assert r.status == "optimal"
    assert abs(r.gap) <= 1e-6 and r.primal_infeas <= 1e-6
    assert r.value > 0


def test_infeasible_problem_is_flagged_by_every_solver_path():
    """a token that must be received in a quantity no pool can deliver (cvxpy: prob.status == 'infeasible'): the python
    loop, the persistent kernel and the per-thread batch solver all end uncertified and the API reports 'infeasible'"""
    class Need:                                   # psi_1 == +100 with 10 in the only pool; objective: psi_0
        def spec(self, n):
            return cf.DualSpec(np.array([1.0, 0.0]), np.array([0.0, -100.0]), np.array([False, True]), np.array([True, False]))
    hp_list = cf.HostPools.from_lists(2, [[0, 1]], [[10.0, 10.0]], [0.997], ["product"], [None])
    hp_pair = cf.HostPools.from_pairs(2, [[0, 1]], [[10.0, 10.0]], [0.997])
    assert cf.solve_pools(hp_list, Need(), method="pools", native=False, max_iter=60).status == "infeasible"
    assert cf.solve_pools(hp_pair, Need(), method="pools", max_iter=60).status == "infeasible"          # persistent kernel
    assert cf.solve_pools(hp_list, Need(), method="thread", max_iter=60).status == "infeasible"
    class Fine(Need):
        def spec(self, n):
            return cf.DualSpec(np.array([1.0, 0.0]), np.array([0.0, -5.0]), np.array([False, True]), np.array([True, False]))
    for kw in (dict(method="pools", native=False), dict(method="thread")):
        r = cf.solve_pools(hp_list, Fine(), **kw)
        assert r.status == "optimal" and abs(r.psi[1] - 5.0) <= 1e-6
    r = cf.solve_pools(hp_pair, Fine(), method="pools")
    assert r.status == "optimal" and abs(r.psi[1] - 5.0) <= 1e-6 and r.info.history == []


def test_error_codes_and_empty_bucket():
    lib = _lib.load()
    b = _lib.Bucket(_lib.KIND_PRODUCT, 2, 10, 10, None, None, None, None, None, None)
    assert lib.cfmm_arb_eval(C.byref(b), 4, None, None, 0.0, None, None, None, None) == -1
    x = torch.zeros(8, **F64)
    b = _lib.Bucket(7, 2, 0, 0, x.data_ptr(), x.data_ptr(), x.data_ptr(), None, None, None)
    assert lib.cfmm_arb_eval(C.byref(b), 4, x.data_ptr(), None, 0.0, x.data_ptr(), x.data_ptr(), None, None) == -2
    b = _lib.Bucket(_lib.KIND_SUM, 3, 1, 1, x.data_ptr(), x.data_ptr(), x.data_ptr(), None, None, None)
    assert lib.cfmm_arb_eval(C.byref(b), 4, x.data_ptr(), None, 0.0, x.data_ptr(), x.data_ptr(), None, None) == -2
    b = _lib.Bucket(_lib.KIND_PRODUCT, 2, 4, 2, x.data_ptr(), x.data_ptr(), x.data_ptr(), None, None, None)
    assert lib.cfmm_arb_eval(C.byref(b), 4, x.data_ptr(), None, 0.0, x.data_ptr(), x.data_ptr(), None, None) == -3
    b = _lib.Bucket(_lib.KIND_PRODUCT, 2, 0, 0, None, None, None, None, None, None)     # empty: a no-op
    assert lib.cfmm_arb_eval(C.byref(b), 4, x.data_ptr(), None, 0.0, x.data_ptr(), x.data_ptr(), None, None) == 0
    with pytest.raises(ValueError):
        cf.HostPools.from_lists(3, [[0, 1, 2]], [[1, 1, 1]], [0.99], ["sum"])
    with pytest.raises(ValueError):
        cf.HostPools.from_pairs(2, [[0, 1]], [[1.0, -1.0]], [0.99]).validate()
    with pytest.raises(ValueError):                  # the device-side validation of the fast path
        cf.PoolStore(cf.HostPools.from_pairs(2, [[0, 1]] * 4, [[1.0, -1.0]] * 4, [0.99] * 4))
    with pytest.raises(ValueError):
        cf.PoolStore(cf.HostPools.from_pairs(2, [[0, 5]] * 4, [[1.0, 1.0]] * 4, [0.99] * 4))
    with pytest.raises(ValueError):
        cf.PoolStore(cf.HostPools.from_pairs(2, [[0, 1]] * 4, [[1.0, 1.0]] * 4, [1.5] * 4))
