"""Generate tests/golden/v3_instance.json: the bounded-liquidity (Uniswap-v3 tick range) extension instance
(cfmm_routing_code_b200/instances.py: v3_instance), solved as a PRIMAL program by scipy SLSQP (oracle/primal_scipy.py).
This pool kind is not in the reference, so these vectors pin the extension against an independent method only.

    python tests/golden/make_golden_v3.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import primal_scipy as PS                      # noqa: E402
from cfmm_routing_code_b200 import instances as I          # noqa: E402

d = I.v3_instance()
args = (3, d["local_indices"], d["reserves"], d["fees"], d["kinds"], d["weights"])
cases = {"arbitrage": (d["market_value"], [("ge", j, 0.0) for j in range(3)])}
for t in (0.0, 5.0, 40.0, 400.0):
    cases[f"swap_0_2_{t:g}"] = ([0, 0, 1.0], [("ge", 0, t), ("ge", 1, 0.0), ("ge", 2, 0.0)])
for t in (1.0, 25.0, 250.0):
    cases[f"swap_1_0_{t:g}"] = ([1.0, 0, 0], [("ge", 0, 0.0), ("ge", 1, t), ("ge", 2, 0.0)])
out = {"how": "scipy SLSQP on the primal program; see make_golden_v3.py"}
for name, (obj, cons) in cases.items():
    r = PS.solve_primal(*args, obj, cons)
    out[name] = dict(value=float(r["value"]), psi=[float(x) for x in r["psi"]],
                     net=[[float(x) for x in (l - dl)] for dl, l in zip(r["deltas"], r["lambdas"])])
    print(name, out[name]["value"])
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "v3_instance.json"), "w") as f:
    json.dump(out, f, indent=1)
