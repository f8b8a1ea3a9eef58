"""Many small routing problems in one launch (needs a B200; there is no CPU fallback).

    python examples/batch_quotes.py

(1) 4096 swap quotes over one pool set -- the loop of two-asset.py:40-100 with 4096 trade sizes instead of 50 -- through
    cf.solve_batch (one problem per GPU thread);
(2) independent markets, each with its own pools and tokens, through cf.solve_many;
(3) a Uniswap-v3 style pool given as its tick ranges (kinds="bounded_product", not a reference atom) next to the
    reference's pool kinds."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfmm_routing_code_b200 as cf                           # noqa: E402
from cfmm_routing_code_b200 import instances as I             # noqa: E402


def main():
    d = I.two_asset_instance()
    hp = cf.HostPools.from_lists(d["n_tokens"], d["local_indices"], d["reserves"], d["fees"], d["kinds"], d["weights"])
    sizes = np.linspace(0.0, 50.0, 4096)
    t0 = time.perf_counter()
    rs = cf.solve_batch(hp, [cf.Swap(d["tok_in"], d["tok_out"], t) for t in sizes], want_trades=False)
    dt = time.perf_counter() - t0
    print(f"4096 quotes in {dt * 1e3:.1f} ms: u(0) = {rs[0].value:.6f}, u(50) = {rs[-1].value:.6f}, "
          f"all optimal: {all(r.status == 'optimal' for r in rs)}")

    a, l = I.arbitrage_instance(), I.liquidation_instance()
    markets = [(cf.HostPools.from_lists(4, a["local_indices"], a["reserves"], a["fees"], a["kinds"], a["weights"]),
                cf.Arbitrage(a["market_value"])),
               (cf.HostPools.from_lists(5, l["local_indices"], l["reserves"], l["fees"], l["kinds"], l["weights"]),
                cf.Liquidate(l["target"], l["current_assets"])),
               (hp, cf.Swap(0, 2, 12.5))]
    for r in cf.solve_many(markets):
        print(f"  market with {len(r.psi)} tokens: value {r.value:.9f} ({r.status}, gap {r.gap:.1e})")

    v = I.v3_instance()
    r = cf.solve(v["local_indices"], v["reserves"], v["fees"], v["kinds"], v["weights"], utility=cf.Swap(0, 2, 40.0))
    print(f"v3-style tick ranges: 40 of token 0 buys {r.value:.6f} of token 2 ({r.status})")
    for i, kind in enumerate(v["kinds"]):
        print(f"  pool {i} ({kind}): net flow {np.round(r.lambdas[i] - r.deltas[i], 6)}")


if __name__ == "__main__":
    main()
