"""Problem instances: the reference's three scripts as data, and the BASELINE.json synthetic configs.

Reference literals (problem DATA, not code):
  arbitrage_instance()   <- arbitrage.py:5-36    (4 tokens, 5 pools, market_value)
  liquidation_instance() <- liquidation.py:5-36  (5 tokens, 5 pools, current_assets)
  two_asset_instance()   <- two-asset.py:7-34    (3 tokens, 5 pools, amounts = linspace(0, 50))
Synthetic generators follow SURVEY.md section 8(d) (cfg 2-5), seeded numpy default_rng.
"""
from __future__ import annotations

import numpy as np


def _five_pools(first_pool, weights0, pair_a, pair_b, pair_c):
    # all three scripts share one shape: a weighted pool over every token, three
    # constant-product pairs and one constant-sum pair (the last two on the same pair)
    return dict(
        local_indices=[list(first_pool), list(pair_a), list(pair_b), list(pair_c), list(pair_c)],
        kinds=["geomean", "product", "product", "product", "sum"],
        weights=[list(weights0), None, None, None, None],
    )


def arbitrage_instance():
    d = _five_pools(range(4), (4, 3, 2, 1), (0, 1), (1, 2), (2, 3))
    d.update(
        n_tokens=4,
        reserves=[[4, 4, 4, 4], [10, 1], [1, 5], [40, 50], [10, 10]],
        fees=[0.998, 0.997, 0.997, 0.997, 0.999],
        market_value=[1.5, 10, 2, 3],
    )
    return d


def liquidation_instance():
    d = _five_pools(range(5), (5, 4, 3, 2, 1), (0, 1), (2, 3), (3, 4))
    d.update(
        n_tokens=5,
        reserves=[[4, 4, 4, 4, 4], [10, 1], [1, 5], [40, 50], [10, 10]],
        fees=[0.998, 0.997, 0.997, 0.997, 0.999],
        current_assets=[2, 1, 3, 5, 10],
        target=4,
    )
    return d


def two_asset_instance():
    d = _five_pools(range(3), (3, 2, 1), (0, 1), (1, 2), (0, 2))
    d.update(
        n_tokens=3,
        reserves=[[3, 0.2, 1], [10, 1], [1, 10], [20, 50], [10, 10]],
        fees=[0.98, 0.99, 0.96, 0.97, 0.99],
        amounts=np.linspace(0, 50),
        tok_in=0,
        tok_out=2,
    )
    return d


# ------------------------------------------------------------------------------------------
# synthetic configs (BASELINE.json configs[1..4])
# ------------------------------------------------------------------------------------------
_FEES = np.array([0.997, 0.999, 0.9995])


def synth_const_product(m, n_tokens, seed, mispricing=0.02):
    """cfg 2 (m=10_000, n=256, seed 0) and cfg 5 (m=1_000_000, n=4096, seed 3)."""
    rng = np.random.default_rng(seed)
    p = np.exp(rng.standard_normal(n_tokens))
    a = rng.integers(0, n_tokens, m)
    b = (a + 1 + rng.integers(0, n_tokens - 1, m)) % n_tokens          # b != a
    liq = np.exp(8.0 + 1.5 * rng.standard_normal(m))
    Ra = liq / p[a] * np.exp(mispricing * rng.standard_normal(m))
    Rb = liq / p[b] * np.exp(mispricing * rng.standard_normal(m))
    gamma = _FEES[rng.integers(0, 3, m)]
    return dict(
        n_tokens=n_tokens, prices=p,
        idx=np.stack([a, b], 1).astype(np.int32), reserves=np.stack([Ra, Rb], 1), gamma=gamma,
    )


def synth_mixed(m, n_tokens, seed, frac_product=0.6, frac_weighted=0.3, mispricing=0.02):
    """cfg 3 / cfg 4 pool population: const-product, weighted (arity 2..8), const-sum pairs
    on tokens whose prices are within 1 %.  Returned in list form (CSR is built by the
    caller) plus the price vector."""
    rng = np.random.default_rng(seed)
    p = np.exp(rng.standard_normal(n_tokens))
    # make near-pegged token pairs exist: every 10th token copies its neighbour's price +-0.5 %
    peg = np.arange(1, n_tokens, 10)
    p[peg] = p[peg - 1] * np.exp(0.005 * rng.standard_normal(len(peg)))
    n_cp = int(round(frac_product * m)); n_w = int(round(frac_weighted * m)); n_cs = m - n_cp - n_w
    ptr = [0]; idx = []; res = []; wts = []; gam = []; kind = []
    cp = synth_const_product(n_cp, n_tokens, seed + 1000, mispricing)
    # rebuild const-product reserves against THIS price vector
    a, b = cp["idx"][:, 0], cp["idx"][:, 1]
    liq = np.exp(8.0 + 1.5 * rng.standard_normal(n_cp))
    Ra = liq / p[a] * np.exp(mispricing * rng.standard_normal(n_cp))
    Rb = liq / p[b] * np.exp(mispricing * rng.standard_normal(n_cp))
    out_idx = [np.stack([a, b], 1)]; out_R = [np.stack([Ra, Rb], 1)]
    out_w = [np.full((n_cp, 2), 0.5)]; out_g = [cp["gamma"]]; out_k = [np.zeros(n_cp, np.uint8)]
    out_ar = [np.full(n_cp, 2)]
    # weighted pools
    ar = rng.integers(2, 9, n_w)
    for k in range(2, 9):
        sel = np.nonzero(ar == k)[0]
        if len(sel) == 0:
            continue
        mk = len(sel)
        toks = np.argsort(rng.random((mk, n_tokens)), axis=1)[:, :k] if n_tokens <= 4096 and mk * n_tokens <= 5e7 \
            else np.stack([rng.choice(n_tokens, k, replace=False) for _ in range(mk)])
        w = rng.dirichlet(np.ones(k), mk)
        w = np.maximum(w, 0.02); w /= w.sum(1, keepdims=True)
        L = np.exp(8.0 + 1.5 * rng.standard_normal(mk))
        R = L[:, None] * w / p[toks] * np.exp(mispricing * rng.standard_normal((mk, k)))
        out_idx.append(toks); out_R.append(R); out_w.append(w)
        out_g.append(_FEES[rng.integers(0, 3, mk)]); out_k.append(np.zeros(mk, np.uint8))
        out_ar.append(np.full(mk, k))
    # const-sum pairs on pegged tokens
    pa = peg[rng.integers(0, len(peg), n_cs)]
    pb = pa - 1
    Rcs = np.exp(6.0 + rng.standard_normal((n_cs, 2)))
    out_idx.append(np.stack([pa, pb], 1)); out_R.append(Rcs); out_w.append(np.zeros((n_cs, 2)))
    out_g.append(_FEES[rng.integers(0, 3, n_cs)]); out_k.append(np.ones(n_cs, np.uint8))
    out_ar.append(np.full(n_cs, 2))
    arity = np.concatenate(out_ar)
    pool_ptr = np.concatenate([[0], np.cumsum(arity)]).astype(np.int64)
    return dict(
        n_tokens=n_tokens, prices=p, pool_ptr=pool_ptr,
        tok_idx=np.concatenate([x.ravel() for x in out_idx]).astype(np.int32),
        reserves=np.concatenate([x.ravel() for x in out_R]),
        weights=np.concatenate([x.ravel() for x in out_w]),
        gamma=np.concatenate(out_g), kind=np.concatenate(out_k),
    )


def synth_basket(n_tokens, prices, seed, n_assets=16, scale=1e-3, liq_mean=np.exp(8.0)):
    """cfg 4 basket: a_j = exp(N(0,1)) * Lbar / p_j * 1e-3 on 16 random tokens, target token 0."""
    rng = np.random.default_rng(seed)
    toks = rng.choice(np.arange(1, n_tokens), n_assets, replace=False)
    a = np.zeros(n_tokens)
    a[toks] = np.exp(rng.standard_normal(n_assets)) * liq_mean / prices[toks] * scale
    return a


# ------------------------------------------------------------------------------------------
# bounded-liquidity constant product (a Uniswap-v3 tick range) -- not in the reference: the first
# "more trading functions" extension behind the per-pool interface
# ------------------------------------------------------------------------------------------
def v3_position(liquidity, p_lo, p_hi, p):
    """Real reserves and virtual-reserve offsets of a concentrated-liquidity position: liquidity L on the price range
    [p_lo, p_hi] (token-1 per token-0) at current price p.  Returns (reserves [x, y], offsets [o_x, o_y]) such that
    (x + o_x)(y + o_y) = L^2 is the position's curve and x, y >= 0 what it can actually pay out."""
    p = float(np.clip(p, p_lo, p_hi))
    sp, sa, sb = np.sqrt(p), np.sqrt(p_lo), np.sqrt(p_hi)
    return [liquidity * (1 / sp - 1 / sb), liquidity * (sp - sa)], [liquidity / sb, liquidity * sa]


def v3_instance():
    """3 tokens; two tick ranges of a v3 pool on (0, 1) (adjacent ranges: the upper one holds token 0 only), one
    in-range position on (0, 2), a constant-product pair (1, 2) and a weighted 3-token pool."""
    r_a, o_a = v3_position(100.0, 0.8, 1.25, 1.0)
    r_b, o_b = v3_position(80.0, 1.25, 1.6, 1.0)        # out of range: all in token 0
    r_c, o_c = v3_position(50.0, 1.9, 2.2, 2.0)
    return dict(
        n_tokens=3,
        local_indices=[[0, 1], [0, 1], [0, 2], [1, 2], [0, 1, 2]],
        reserves=[r_a, r_b, r_c, [30.0, 20.0], [10.0, 12.0, 8.0]],
        fees=[0.997, 0.997, 0.9995, 0.997, 0.99],
        kinds=["bounded_product", "bounded_product", "bounded_product", "product", "geomean"],
        weights=[o_a, o_b, o_c, None, [2, 1, 1]],
        market_value=[1.3, 1.0, 0.45],
    )
