"""Shared builders for tests (not a test module)."""
import numpy as np

import cfmm_routing_code_b200 as cf
from cfmm_routing_code_b200 import instances as I
from oracle import cfmm_oracle as O


def host_pools(d):
    return cf.HostPools.from_lists(d["n_tokens"], d["local_indices"], d["reserves"], d["fees"], d["kinds"],
                                   d["weights"])


def oracle_pools(hp):
    return O.Pools(hp.n_tokens, hp.pool_ptr, hp.tok_idx, hp.reserves, hp.weights, hp.gamma, hp.kind)


def cp_host_pools(m, n, seed):
    s = I.synth_const_product(m, n, seed)
    return cf.HostPools.from_pairs(n, s["idx"], s["reserves"], s["gamma"]), s


def mixed_host_pools(m, n, seed):
    s = I.synth_mixed(m, n, seed)
    hp = cf.HostPools(n, s["pool_ptr"], s["tok_idx"], s["reserves"], s["weights"], s["gamma"], s["kind"])
    return hp, s


def random_prices(prices, seed, spread=0.05):
    rng = np.random.default_rng(seed)
    return prices * np.exp(spread * rng.standard_normal(len(prices)))


def random_small_problem(rng, all_kinds=True):
    """A random routing problem of the reference's scale: 3-6 tokens, up to 13 pools of every kind.  A chain of
    constant-product pools over all tokens comes first, so every utility below is feasible.  Returns (HostPools,
    list-form dict, token prices)."""
    n = int(rng.integers(3, 7)); m = int(rng.integers(n, 14))
    prices = np.exp(rng.normal(0, 1, n))
    li, res, fees, kinds, w = [], [], [], [], []
    pool_kinds = ["product", "geomean", "sum", "bounded_product"] if all_kinds else ["product", "geomean", "sum"]
    probs = [0.4, 0.25, 0.15, 0.2] if all_kinds else [0.5, 0.3, 0.2]
    for i in range(m):
        kd = str(rng.choice(pool_kinds, p=probs))
        k = 2 if kd != "geomean" else int(rng.integers(2, min(n, 5) + 1))
        toks = rng.choice(n, k, replace=False)
        if i < n - 1:
            kd, k, toks = "product", 2, np.array([i, i + 1])
        liq = np.exp(rng.normal(4, 1.0))
        noise = np.exp(0.05 * rng.standard_normal(k))
        if kd == "geomean":
            ww = rng.dirichlet(np.ones(k)); R = liq * ww / prices[toks] * noise; w.append(list(ww))
        elif kd == "bounded_product":
            p = prices[toks[0]] / prices[toks[1]]
            lo, hi = p * np.exp(-rng.uniform(0.01, 0.3)), p * np.exp(rng.uniform(0.01, 0.3))
            R, o = I.v3_position(liq / np.sqrt(prices[toks[0]] * prices[toks[1]]), lo, hi,
                                 p * np.exp(0.02 * rng.standard_normal()))
            w.append(o)
        else:
            R = liq / prices[toks] * noise; w.append(None)
        li.append([int(t) for t in toks]); res.append([float(x) for x in R]); kinds.append(kd)
        fees.append(float(rng.choice([0.997, 0.999, 0.9995])))
    d = dict(n_tokens=n, local_indices=li, reserves=res, fees=fees, kinds=kinds, weights=w)
    return host_pools(d), d, prices


def random_utilities(rng, n, prices):
    """one of each utility of the reference (arbitrage.py:57,77 / two-asset.py:66,86 / liquidation.py:57,77-80)"""
    us = [O.Utility.arbitrage(prices * np.exp(0.03 * rng.standard_normal(n)))]
    i, o = rng.choice(n, 2, replace=False)
    us.append(O.Utility.swap(n, int(i), int(o), float(np.exp(rng.normal(2, 1.5)) / prices[i])))
    basket = np.zeros(n); tgt = int(rng.integers(n))
    for j in rng.choice(n, 2, replace=False):
        if j != tgt:
            basket[j] = float(np.exp(rng.normal(1, 1)) / prices[j])
    us.append(O.Utility.liquidate(n, tgt, basket))
    return us


def check_blocked_tables(t, idx, order, n, P, rs, ts, cap):
    """Invariants of a blocked layout (tables `t` as CPU tensors, idx (2, m) int64, order = pool at each blocked position):
    emulate the kernels' scatter -- the pool phase writes the two flows of every pool to its slots of the tile's
    row-ordered array, one thread sums each row -- and compare with a plain index_add; rows longest first, inside the
    tile's 2 P slots; local ids map back to the pools' tokens."""
    import torch
    M, T = t["M"], t["n_tiles"]
    desc = t["desc"].to(torch.int64)
    assert M == T * P
    nq = len(order)
    q = torch.arange(nq)
    tl = q // P
    f = torch.randn(nq, 2, dtype=torch.float64)                     # flows of (pool, slot), blocked order
    pos = t["pos"].to(torch.int64)[:nq] & 0xffffffff
    g = torch.zeros(T, 2 * P, dtype=torch.float64)                  # the pool phase scatters into row order
    g[tl, pos & 0xffff] = f[:, 0]
    g[tl, pos >> 16] = f[:, 1]
    padpos = t["pos"].to(torch.int64)[nq:] & 0xffffffff             # padding pools of the last tile: slots past the real flows
    assert bool(((padpos & 0xffff) < 2 * P).all() and ((padpos >> 16) < 2 * P).all())
    rows = t["rows"].to(torch.int64) & 0xffffffff
    out = torch.zeros(n, dtype=torch.float64)
    for tile in range(T):
        ntok, nrow = desc[tile, 0].item(), desc[tile, 1].item()
        assert 0 < ntok <= ts and 0 < nrow <= rs
        w = rows[tile, :nrow]
        st, ln, lt = w & 0xffff, (w >> 16) & 0x3f, w >> 22
        assert bool((ln[:-1] >= ln[1:]).all()) and int(ln.max()) <= cap and int(ln.min()) >= 1     # longest rows first
        assert int((st + ln).max()) <= 2 * P                        # rows stay inside the tile's 2 P flow slots
        assert bool((lt < ntok).all())
        for r in range(nrow):
            out[t["tok"][tile, lt[r]]] += g[tile, st[r]:st[r] + ln[r]].sum()
    a, b = idx[0][order], idx[1][order]
    ref = torch.zeros(n, dtype=torch.float64)
    ref.index_add_(0, a, f[q, 0]); ref.index_add_(0, b, f[q, 1])
    assert float((out - ref).abs().max()) <= 1e-12 * float(ref.abs().max())
    lid = t["lid"].to(torch.int64)[:nq] & 0xffffffff
    assert bool((t["tok"][tl, lid & 0xffff] == a).all() and (t["tok"][tl, lid >> 16] == b).all())
