"""Generate tests/golden/reference_instances.json  (run HERE, in the build container; commit the output).

The reference cannot run in this image (cvxpy absent), so the golden vectors come from
oracle/primal_scipy.py: the reference's primal program (arbitrage.py:50-82, liquidation.py:50-85,
two-asset.py:59-91) solved by scipy SLSQP -- independent of the dual decomposition under test.
SURVEY.md section 8c lists the same numbers, derived a third way (zero duality gap).

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import primal_scipy as PS                      # noqa: E402
from cfmm_routing_code_b200 import instances as I          # noqa: E402


def pack(r):
    return dict(value=float(r["value"]), psi=[float(x) for x in r["psi"]],
                deltas=[[float(x) for x in d] for d in r["deltas"]],
                lambdas=[[float(x) for x in d] for d in r["lambdas"]])


out = {"how": "scipy SLSQP on the reference's primal program; see make_golden.py",
       "survey_8c": {"arbitrage": 21.49980876354458, "liquidation": 15.883010841078224,
                     "two_asset_t0": 6.23300013143708, "two_asset_t50": 44.18202040136949}}
d = I.arbitrage_instance()
out["arbitrage"] = pack(PS.solve_primal(4, d["local_indices"], d["reserves"], d["fees"], d["kinds"],
                                        d["weights"], d["market_value"], [("ge", j, 0.0) for j in range(4)]))
d = I.liquidation_instance()
c = np.zeros(5); c[4] = 1
out["liquidation"] = pack(PS.solve_primal(5, d["local_indices"], d["reserves"], d["fees"], d["kinds"],
                                          d["weights"], c, [("eq", j, d["current_assets"][j]) for j in range(4)]))
d = I.two_asset_instance()
c = np.zeros(3); c[2] = 1
sweep = []
for t in d["amounts"]:
    r = PS.solve_primal(3, d["local_indices"], d["reserves"], d["fees"], d["kinds"], d["weights"], c,
                        [("ge", 0, float(t)), ("ge", 1, 0.0), ("ge", 2, 0.0)])
    p = pack(r); p["t"] = float(t)
    sweep.append(p)
out["two_asset"] = sweep
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_instances.json"), "w") as f:
    json.dump(out, f, indent=1)
print("arbitrage", out["arbitrage"]["value"], "liquidation", out["liquidation"]["value"],
      "two-asset", sweep[0]["value"], sweep[-1]["value"])
