"""Batched small-problem solver (csrc/cfmm_small.cuh / cfmm_small.cu): one whole prob.solve() per GPU thread.

CPU part: the per-thread solver, compiled for the host by tests/small_host.py, against oracle/cfmm_oracle.py::solve --
same algorithm, so iterates agree to rounding and the evaluation counts are equal.  GPU part: the kernel through the
C ABI against the oracle and the golden vectors."""
import numpy as np
import pytest

import cfmm_routing_code_b200 as cf
from cfmm_routing_code_b200 import instances as I
from oracle import cfmm_oracle as O
import helpers as H
import small_host


def _reference_cases():
    d = I.arbitrage_instance(); hp = H.host_pools(d)
    yield "arbitrage", hp, [O.Utility.arbitrage(d["market_value"])]
    d = I.liquidation_instance(); hp = H.host_pools(d)
    yield "liquidation", hp, [O.Utility.liquidate(hp.n_tokens, d["target"], d["current_assets"])]
    d = I.two_asset_instance(); hp = H.host_pools(d)
    yield "two_asset", hp, [O.Utility.swap(hp.n_tokens, d["tok_in"], d["tok_out"], t) for t in d["amounts"]]


def _small_mixed(seed, m=24, n=9):
    s = I.synth_mixed(m, n, seed)
    hp = cf.HostPools(n, s["pool_ptr"], s["tok_idx"], s["reserves"], s["weights"], s["gamma"], s["kind"])
    return hp, s


def _check_against_oracle(hp, specs, out, tol=1e-9, rtol=1e-9):
    op = H.oracle_pools(hp)
    for p, u in enumerate(specs):
        r = O.solve(op, u, tol=tol)
        st = out["stats"][p]
        assert int(st[7]) == {"optimal": 0, "max_iter": 1, "stalled": 2}[r.status], p
        scale = max(abs(r.dual_value), 1e-300)
        assert abs(st[0] - r.value) <= rtol * scale, (p, st[0], r.value)
        assert abs(st[1] - r.dual_value) <= rtol * scale
        np.testing.assert_allclose(out["nu"][p], r.nu, rtol=1e-7)
        gross = np.abs(np.concatenate(r.deltas)).sum() + np.abs(np.concatenate(r.lambdas)).sum()
        np.testing.assert_allclose(out["psi"][p], r.psi, atol=1e-7 * max(gross, 1e-300))
        if out.get("delta") is not None and out["delta"].shape[0] == len(specs):
            np.testing.assert_allclose(out["delta"][p], np.concatenate(r.deltas), atol=1e-6 * max(gross, 1e-300))
            np.testing.assert_allclose(out["lam"][p], np.concatenate(r.lambdas), atol=1e-6 * max(gross, 1e-300))


# ------------------------------------------------------------------------------------------------- CPU (host build)
@pytest.mark.parametrize("interleave", [0, 1])
def test_host_build_reference_instances_match_oracle_step_for_step(interleave, golden):
    for name, hp, specs in _reference_cases():
        out = small_host.solve(hp, specs, tol=1e-9, interleave=interleave)
        _check_against_oracle(hp, specs, out)
        op = H.oracle_pools(hp)
        for p in (0, len(specs) - 1):           # same control flow => same number of evaluations and iterations
            r = O.solve(op, specs[p], tol=1e-9)
            assert (int(out["stats"][p][5]), int(out["stats"][p][6])) == (r.iters, r.evals), name
    assert abs(out["stats"][0][0] - golden["survey_8c"]["two_asset_t0"]) <= 1e-8 * 6.3
    assert abs(out["stats"][49][0] - golden["survey_8c"]["two_asset_t50"]) <= 1e-8 * 44.2


def test_host_build_matches_golden_with_the_end_to_end_tolerances(golden):
    """the assertions of test_gpu_parity.py::test_reference_instances_end_to_end, on the host build"""
    cases = {name: (hp, specs) for name, hp, specs in _reference_cases()}
    for name in ("arbitrage", "liquidation"):
        hp, specs = cases[name]
        out = small_host.solve(hp, specs, tol=1e-9)
        g = golden[name]
        assert int(out["stats"][0][7]) == 0
        val = out["stats"][0][0]
        assert abs(val - g["value"]) <= 1e-6 * abs(g["value"])
        assert abs(val - golden["survey_8c"][name]) <= 1e-8 * abs(val)
        np.testing.assert_allclose(out["psi"][0], g["psi"], atol=1e-6 * (3.25 if name == "arbitrage" else 15.9))
        ptr = hp.pool_ptr
        for i in range(hp.m):
            np.testing.assert_allclose(out["delta"][0][ptr[i]:ptr[i + 1]], g["deltas"][i], atol=5e-5)
            np.testing.assert_allclose(out["lam"][0][ptr[i]:ptr[i + 1]], g["lambdas"][i], atol=5e-5)


def test_host_build_two_asset_sweep_matches_golden(golden):
    _, hp, specs = list(_reference_cases())[2]
    out = small_host.solve(hp, specs, tol=1e-9)
    for j in range(50):
        g = golden["two_asset"][j]
        assert abs(out["stats"][j][0] - g["value"]) <= 1e-6 * max(1.0, g["value"]), j
        np.testing.assert_allclose(out["psi"][j], g["psi"], atol=2e-5)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_host_build_mixed_problems_match_oracle(seed):
    hp, s = _small_mixed(seed)
    rng = np.random.default_rng(seed)
    specs = [O.Utility.arbitrage(s["prices"] * np.exp(0.05 * rng.standard_normal(hp.n_tokens))) for _ in range(3)]
    basket = np.zeros(hp.n_tokens); basket[1:4] = rng.uniform(0.1, 1.0, 3) * 1e-2 / s["prices"][1:4] * 3000.0
    specs.append(O.Utility.liquidate(hp.n_tokens, 0, basket))
    specs.append(O.Utility.swap(hp.n_tokens, 2, 5, 10.0 / s["prices"][2]))
    out = small_host.solve(hp, specs, tol=1e-9)
    _check_against_oracle(hp, specs, out, rtol=1e-8)


def test_host_build_disjoint_pool_ranges_and_rejected_input():
    """problems over different pool subsets of one CSR array; a pool the closed forms do not cover is refused"""
    hp, s = _small_mixed(5, m=30, n=8)
    ranges = np.array([[0, 10], [10, 30], [0, 30]], np.int64)
    u = O.Utility.arbitrage(s["prices"])
    out = small_host.solve(hp, [u, u, u], tol=1e-9, pool_range=ranges)
    for p, (lo, hi) in enumerate(ranges):
        sub = cf.HostPools(hp.n_tokens, hp.pool_ptr[lo:hi + 1] - hp.pool_ptr[lo],
                           hp.tok_idx[hp.pool_ptr[lo]:hp.pool_ptr[hi]], hp.reserves[hp.pool_ptr[lo]:hp.pool_ptr[hi]],
                           hp.weights[hp.pool_ptr[lo]:hp.pool_ptr[hi]], hp.gamma[lo:hi], hp.kind[lo:hi])
        r = O.solve(H.oracle_pools(sub), u, tol=1e-9)
        assert abs(out["stats"][p][0] - r.value) <= 1e-8 * max(abs(r.dual_value), 1e-300)
    d = I.arbitrage_instance(); bad = H.host_pools(d)
    bad.tok_idx = bad.tok_idx.copy(); bad.tok_idx[1] = 7          # token index outside n_tokens
    out = small_host.solve(bad, [O.Utility.arbitrage(d["market_value"])])
    assert int(out["stats"][0][7]) == 3 and np.isnan(out["stats"][0][0])


def _many_problems():
    probs = []
    d = I.arbitrage_instance(); probs.append((H.host_pools(d), cf.Arbitrage(d["market_value"])))
    d = I.liquidation_instance(); probs.append((H.host_pools(d), cf.Liquidate(d["target"], d["current_assets"])))
    d = I.two_asset_instance(); probs.append((H.host_pools(d), cf.Swap(d["tok_in"], d["tok_out"], 12.5)))
    hp, s = _small_mixed(11, m=20, n=9); probs.append((hp, cf.Arbitrage(s["prices"])))
    return probs


def _oracle_for(hp, util):
    sp = util.spec(hp.n_tokens)
    return O.solve(H.oracle_pools(hp), O.Utility(sp.c, sp.a, sp.eq, sp.pinned), tol=1e-9)


def test_packing_of_independent_problems_on_the_host_build():
    """batch.pack_problems (what solve_many launches): problems with different pools AND different token counts"""
    from cfmm_routing_code_b200 import batch as B
    probs = _many_problems()
    merged, ranges, c, a, fl, nu, nnz_max = B.pack_problems(probs)
    assert merged.n_tokens == 9 and len(merged.gamma) == 5 + 5 + 5 + 20
    assert nnz_max == max(int(hp.pool_ptr[-1]) for hp, _ in probs) and ranges[-1, 1] == 35
    out = small_host.solve_raw(merged, c, a, fl, nu, pool_range=ranges, tol=1e-9)
    off = 0
    for p, (hp, u) in enumerate(probs):
        r = _oracle_for(hp, u)
        assert int(out["stats"][p][7]) == 0
        assert abs(out["stats"][p][0] - r.value) <= 1e-8 * max(abs(r.dual_value), 1e-300)
        np.testing.assert_allclose(out["psi"][p, :hp.n_tokens], r.psi, atol=1e-6 * max(1.0, np.abs(r.psi).max()))
        assert np.all(out["psi"][p, hp.n_tokens:] == 0.0) and np.all(out["nu"][p, hp.n_tokens:] == 1.0)
        nnz = int(hp.pool_ptr[-1])
        np.testing.assert_allclose(out["lam"][0, off:off + nnz], np.concatenate(r.lambdas), atol=1e-5 * max(1.0, np.abs(r.psi).max()))
        off += nnz


def _v3_cases():
    d = I.v3_instance(); hp = H.host_pools(d)
    specs = [O.Utility.arbitrage(d["market_value"])]
    specs += [O.Utility.swap(3, 0, 2, t) for t in (0.0, 5.0, 40.0, 400.0)]      # 400: every range on the way is drained
    specs += [O.Utility.swap(3, 1, 0, t) for t in (1.0, 25.0, 250.0)]
    specs.append(O.Utility.liquidate(3, 2, [3.0, 7.0, 0.0]))
    return hp, specs


def _v3_golden():
    import json, os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "v3_instance.json")) as f:
        g = json.load(f)
    names = ["arbitrage"] + [f"swap_0_2_{t}" for t in (0, 5, 40, 400)] + [f"swap_1_0_{t}" for t in (1, 25, 250)]
    return [g[k] for k in names]            # same order as the first 8 of _v3_cases()


def test_host_build_bounded_product_pools_match_oracle():
    """the v3-style pools (kind 3) in the per-thread solver: in range, at the cap, out of range"""
    hp, specs = _v3_cases()
    out = small_host.solve(hp, specs, tol=1e-9)
    _check_against_oracle(hp, specs, out)
    assert np.all(out["stats"][:, 7] == 0)
    for p, g in enumerate(_v3_golden()):                 # SLSQP on the primal program (tests/golden/make_golden_v3.py)
        assert abs(out["stats"][p][0] - g["value"]) <= 1e-6 * max(1.0, abs(g["value"])), p
        np.testing.assert_allclose(out["psi"][p], g["psi"], atol=2e-5)
    # drained ranges pay out exactly their real reserves
    lam = out["lam"][4]                                  # swap 400 of token 0 for token 2
    np.testing.assert_allclose(lam[hp.pool_ptr[2] + 1], hp.reserves[hp.pool_ptr[2] + 1], rtol=1e-12)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_host_build_random_problems_of_every_kind_and_utility(seed):
    """90 random problems per seed (all four pool kinds, three utilities) against the oracle; every problem ends with a certificate (gap and infeasibility <= 1e-7) even where the KKT residual
    stalls a hair above the requested 1e-8 at the fp64 floor of the constant-sum ramp"""
    rng = np.random.default_rng(seed)
    n_opt = 0
    for _ in range(30):
        hp, d, prices = H.random_small_problem(rng)
        specs = H.random_utilities(rng, hp.n_tokens, prices)
        out = small_host.solve(hp, specs, tol=1e-8)
        op = H.oracle_pools(hp)
        for p, u in enumerate(specs):
            r = O.solve(op, u, tol=1e-8)
            st = out["stats"][p]
            assert abs(st[0] - r.value) <= 1e-7 * max(abs(r.dual_value), 1e-300)
            assert abs(st[2]) <= 1e-7 and st[3] <= 1e-7, (st[2], st[3])
            n_opt += int(st[7]) == 0
    assert n_opt >= 88


def test_host_build_random_problems_against_the_slsqp_primal():
    """the independent pin: four random problems re-solved as primal programs by scipy SLSQP"""
    from oracle import primal_scipy as PS
    rng = np.random.default_rng(12)
    for _ in range(4):
        hp, d, prices = H.random_small_problem(rng)
        n = hp.n_tokens
        u = H.random_utilities(rng, n, prices)[1]                  # the swap: always feasible, bounded
        out = small_host.solve(hp, [u], tol=1e-9)
        cons = [("ge", j, float(u.a[j])) for j in range(n)]
        pr = PS.solve_primal(n, d["local_indices"], d["reserves"], d["fees"], d["kinds"], d["weights"], u.c, cons)
        assert int(out["stats"][0][7]) == 0
        assert abs(out["stats"][0][0] - pr["value"]) <= 2e-6 * max(1.0, abs(pr["value"]))


def test_host_build_infeasible_liquidation_is_not_reported_optimal():
    """a basket token no pool can route to the target (liquidation.py:77-80 would be infeasible in cvxpy)"""
    hp = cf.HostPools.from_lists(3, [[0, 1]], [[10.0, 10.0]], [0.997], ["product"], [None])
    u = O.Utility.liquidate(3, 0, [0.0, 1.0, 2.0])                # token 2 is in no pool
    out = small_host.solve(hp, [u], tol=1e-8)
    assert int(out["stats"][0][7]) != 0                          # the price of token 2 just drifts towards its floor
    assert abs(out["psi"][0][2] + 2.0) > 1.0                      # and the basket constraint is not met
    with pytest.raises(ValueError, match="infeasible"):          # the API refuses such problems before any launch
        cf.solve_pools(hp, cf.Liquidate(0, [0.0, 1.0, 2.0]))
    with pytest.raises(ValueError, match="infeasible"):
        cf.solve_batch(hp, [cf.Liquidate(0, [0.0, 1.0, 2.0])])


class _HostStore:
    """stand-in for batch.CsrStore in the plumbing test below: keeps the HostPools, owns no device memory"""
    def __init__(self, hp, device="cuda"):
        import torch
        self.hp, self.device = hp, torch.device("cpu")
        self.n_tokens, self.m, self.nnz = hp.n_tokens, hp.m, int(len(hp.tok_idx))


def _host_solve_batch_device(store, c, a, flags, nu, tol=1e-8, want_trades=True, pool_range=None, max_outer=60,
                             max_inner=100, nnz_max=0, lanes=None):
    import torch
    out = small_host.solve_raw(store.hp, c.numpy(), a.numpy(), flags.numpy(), nu.numpy(),
                               None if pool_range is None else pool_range.numpy(), tol=tol)
    nu.copy_(torch.as_tensor(out["nu"]))
    t = torch.as_tensor
    return t(out["psi"]), t(out["stats"]), (t(out["delta"]) if want_trades else None), (t(out["lam"]) if want_trades else None)


def test_python_plumbing_of_batch_entry_points_with_the_host_build(monkeypatch, golden):
    """solve(method='thread'), solve_sweep(batched), solve_batch and solve_many: packing, routing and unpacking of the
    results, with the kernel launch replaced by the host build of the same solver (no GPU here)"""
    from cfmm_routing_code_b200 import batch as B
    monkeypatch.setattr(B, "CsrStore", _HostStore)
    monkeypatch.setattr(B, "solve_batch_device", _host_solve_batch_device)
    d = I.arbitrage_instance()
    r = cf.solve(d["local_indices"], d["reserves"], d["fees"], d["kinds"], d["weights"],
                 utility=cf.Arbitrage(d["market_value"]), tol=1e-9, method="thread")
    assert r.status == "optimal" and abs(r.value - golden["survey_8c"]["arbitrage"]) <= 1e-8 * 21.5
    assert len(r.deltas) == 5 and r.deltas[0].shape == (4,) and r.info is None
    np.testing.assert_allclose(r.lambdas[4], golden["arbitrage"]["lambdas"][4], atol=5e-5)
    d = I.two_asset_instance()
    rs = cf.solve_sweep(d["local_indices"], d["reserves"], d["fees"], d["kinds"], d["weights"],
                        [cf.Swap(d["tok_in"], d["tok_out"], t) for t in d["amounts"]], tol=1e-9)
    assert len(rs) == 50 and all(x.status == "optimal" for x in rs)
    for j, x in enumerate(rs):
        assert abs(x.value - golden["two_asset"][j]["value"]) <= 1e-6 * max(1.0, golden["two_asset"][j]["value"])
    probs = _many_problems()
    out = cf.solve_many(probs, tol=1e-9)
    for (hp, u), x in zip(probs, out):
        ro = _oracle_for(hp, u)
        assert x.status == "optimal" and x.psi.shape == (hp.n_tokens,) and len(x.lambdas) == hp.m
        assert abs(x.value - ro.value) <= 1e-8 * max(abs(ro.dual_value), 1e-300)
        for i in range(hp.m):
            np.testing.assert_allclose(x.lambdas[i], ro.lambdas[i], atol=1e-5 * max(1.0, np.abs(ro.psi).max()))
    assert cf.solve_batch(probs[0][0], []) == [] and cf.solve_many([]) == []


# ------------------------------------------------------------------------------------------------- GPU (the product)
def _to_api(u):
    class _U:
        def spec(self, n):
            return cf.DualSpec(u.c, u.a, u.eq, u.pinned)
    return _U()


@pytest.mark.gpu
def test_batch_kernel_reference_instances_match_oracle_and_golden(golden):
    for name, hp, specs in _reference_cases():
        rs = cf.solve_batch(hp, [_to_api(u) for u in specs], tol=1e-9)
        out = dict(stats=np.array([[r.value, r.dual_value, r.gap, r.primal_infeas, 0, r.iters, r.evals,
                                    {"optimal": 0, "max_iter": 1, "stalled": 2}[r.status]] for r in rs]),
                   nu=np.stack([r.nu for r in rs]), psi=np.stack([r.psi for r in rs]),
                   delta=np.stack([np.concatenate(r.deltas) for r in rs]),
                   lam=np.stack([np.concatenate(r.lambdas) for r in rs]))
        _check_against_oracle(hp, specs, out)
        if name == "two_asset":
            for j, r in enumerate(rs):
                g = golden["two_asset"][j]
                assert r.status == "optimal" and abs(r.value - g["value"]) <= 1e-6 * max(1.0, g["value"]), j
        else:
            assert abs(rs[0].value - golden["survey_8c"][name]) <= 1e-8 * abs(rs[0].value)


@pytest.mark.gpu
def test_batch_kernel_agrees_with_its_host_build():
    """4096 random swap quotes over one mixed pool set, against the host build of the same source"""
    hp, s = _small_mixed(7, m=40, n=12)
    rng = np.random.default_rng(7)
    B = 4096
    specs = []
    for _ in range(B):
        i, o = rng.choice(hp.n_tokens, 2, replace=False)
        specs.append(O.Utility.swap(hp.n_tokens, int(i), int(o), float(rng.uniform(0.0, 50.0) / s["prices"][i])))
    rs = cf.solve_batch(hp, [_to_api(u) for u in specs], tol=1e-9, want_trades=False)
    ref = small_host.solve(hp, specs, tol=1e-9)
    val = np.array([r.value for r in rs]); dual = np.array([r.dual_value for r in rs])
    np.testing.assert_allclose(val, ref["stats"][:, 0], rtol=1e-8, atol=1e-9 * np.abs(dual).max())
    status = np.array([{"optimal": 0, "max_iter": 1, "stalled": 2}[r.status] for r in rs])
    assert np.array_equal(status, ref["stats"][:, 7].astype(int))
    # fused multiply-adds on the device round differently, which flips line-search branches on some problems: the
    # evaluation counts agree for most problems (87 % measured) and on average, not one by one
    ev_gpu, ev_host = np.array([r.evals for r in rs], float), ref["stats"][:, 6]
    assert np.mean(ev_gpu == ev_host) >= 0.6
    assert abs(ev_gpu.mean() - ev_host.mean()) <= 0.1 * ev_host.mean()
    assert np.all(np.array([r.gap for r in rs])[status == 0] <= 1e-8)


@pytest.mark.gpu
def test_batched_and_sequential_sweeps_agree():
    d = I.two_asset_instance()
    us = [cf.Swap(d["tok_in"], d["tok_out"], t) for t in d["amounts"][::7]]
    args = (d["local_indices"], d["reserves"], d["fees"], d["kinds"], d["weights"], us)
    a = cf.solve_sweep(*args, tol=1e-9, batched=True)
    b = cf.solve_sweep(*args, tol=1e-9, batched=False)
    for x, y in zip(a, b):
        assert abs(x.value - y.value) <= 1e-7 * max(1.0, abs(y.value))
        np.testing.assert_allclose(x.psi, y.psi, atol=2e-5)


@pytest.mark.gpu
def test_batch_kernel_rejects_what_it_does_not_cover():
    d = I.arbitrage_instance(); hp = H.host_pools(d)
    store = cf.CsrStore(hp)
    store.tok[1] = 7                                            # corrupt a token index on the device
    rs = cf.solve_batch(hp, [cf.Arbitrage(d["market_value"])], store=store)
    assert rs[0].status == "rejected" and np.isnan(rs[0].value)
    big = cf.HostPools.from_pairs(100, np.array([[0, 99]]), np.array([[1.0, 2.0]]), np.array([0.997]))
    with pytest.raises(ValueError):
        cf.CsrStore(big)


@pytest.mark.gpu
def test_solve_many_independent_problems_in_one_launch():
    probs = _many_problems() * 40            # 160 problems, four shapes
    rs = cf.solve_many(probs, tol=1e-9)
    assert len(rs) == 160
    refs = [_oracle_for(hp, u) for hp, u in probs[:4]]
    for p, r in enumerate(rs):
        ro, hp = refs[p % 4], probs[p][0]
        assert r.status == "optimal", p
        assert abs(r.value - ro.value) <= 1e-8 * max(abs(ro.dual_value), 1e-300)
        assert r.psi.shape == (hp.n_tokens,) and len(r.deltas) == hp.m
        np.testing.assert_allclose(r.psi, ro.psi, atol=1e-6 * max(1.0, np.abs(ro.psi).max()))
        for i in range(hp.m):
            np.testing.assert_allclose(r.lambdas[i], ro.lambdas[i], atol=1e-5 * max(1.0, np.abs(ro.psi).max()))


@pytest.mark.gpu
def test_batch_kernel_bounded_product_pools_match_oracle():
    hp, specs = _v3_cases()
    rs = cf.solve_batch(hp, [_to_api(u) for u in specs], tol=1e-9)
    op = H.oracle_pools(hp)
    for u, r in zip(specs, rs):
        ro = O.solve(op, u, tol=1e-9)
        assert r.status == "optimal"
        assert abs(r.value - ro.value) <= 1e-8 * max(abs(ro.dual_value), 1e-300)
        np.testing.assert_allclose(r.psi, ro.psi, atol=1e-6 * max(1.0, np.abs(ro.psi).max()))
    for r, g in zip(rs, _v3_golden()):
        assert abs(r.value - g["value"]) <= 1e-6 * max(1.0, abs(g["value"]))
        np.testing.assert_allclose(r.psi, g["psi"], atol=2e-5)
    for u, g in zip(specs[:3], _v3_golden()):            # and the same pools through the pool-parallel kernels
        r = cf.solve_pools(hp, _to_api(u), method="pools", tol=1e-9)
        assert r.status == "optimal"
        assert abs(r.value - g["value"]) <= 1e-6 * max(1.0, abs(g["value"]))
        np.testing.assert_allclose(r.psi, g["psi"], atol=2e-5)


@pytest.mark.gpu
def test_warp_per_problem_variant_matches_thread_per_problem():
    import torch
    from cfmm_routing_code_b200 import batch as B
    d = I.two_asset_instance(); hp = H.host_pools(d)
    store = cf.CsrStore(hp)
    us = [cf.Swap(d["tok_in"], d["tok_out"], t) for t in d["amounts"]]
    c, a, fl, nu0 = B.pack_utilities(us, hp.n_tokens)
    up = lambda x: torch.as_tensor(x, device="cuda")
    res = {}
    try:
        for lanes in (1, 32):
            nu = up(nu0.copy())
            psi, stats, _, _ = B.solve_batch_device(store, up(c), up(a), up(fl), nu, tol=1e-9, want_trades=False,
                                                    lanes=lanes)
            res[lanes] = (stats.cpu().numpy(), psi.cpu().numpy())
    finally:
        store.lib.cfmm_set_batch_lanes(1)
    s1, s32 = res[1][0], res[32][0]
    assert np.all(s32[:, 7] == 0)
    np.testing.assert_allclose(s32[:, 0], s1[:, 0], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(res[32][1], res[1][1], atol=1e-6)
