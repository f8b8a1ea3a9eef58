"""The reference's three scripts, run through the B200 path: same literals in, same printed lines out.

    python examples/run_reference_instances.py            (needs a B200; there is no CPU fallback)

Prints what arbitrage.py:84, liquidation.py:87 and two-asset.py:96-98 print; writes the two-asset sweep that the
reference plots (two-asset.py:102-118) to output/two_asset_sweep.csv instead of PDFs (matplotlib/TeX are absent)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfmm_routing_code_b200 as cf                           # noqa: E402
from cfmm_routing_code_b200 import instances as I             # noqa: E402


def main():
    d = I.arbitrage_instance()
    r = cf.solve(d["local_indices"], d["reserves"], d["fees"], d["kinds"], d["weights"],
                 utility=cf.Arbitrage(d["market_value"]))
    print(f"Total output value: {r.value}")
    d = I.liquidation_instance()
    r = cf.solve(d["local_indices"], d["reserves"], d["fees"], d["kinds"], d["weights"],
                 utility=cf.Liquidate(d["target"], d["current_assets"]))
    print(f"Total liquidated value: {r.psi[d['target']]}")
    d = I.two_asset_instance()
    # the loop of two-asset.py:40-100 as ONE kernel launch: all 50 trade sizes solved at once, one per GPU thread
    rs = cf.solve_sweep(d["local_indices"], d["reserves"], d["fees"], d["kinds"], d["weights"],
                        [cf.Swap(d["tok_in"], d["tok_out"], t) for t in d["amounts"]])
    rows = []
    for t, r in zip(d["amounts"], rs):
        print(f"Total liquidated value: {r.psi[d['tok_out']]}")
        for i in range(5):
            print(f"Market {i}, delta: {r.deltas[i]}, lambda: {r.lambdas[i]}")
        rows.append([t, r.value] + [x for k in range(5) for x in (r.lambdas[k] - r.deltas[k])])
    os.makedirs("output", exist_ok=True)
    np.savetxt("output/two_asset_sweep.csv", np.asarray(rows), delimiter=",",
               header="t,u_t," + ",".join(f"net_{k}_{j}" for k, l in enumerate(d["local_indices"]) for j in range(len(l))))


if __name__ == "__main__":
    main()
