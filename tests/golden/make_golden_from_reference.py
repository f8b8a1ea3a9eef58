"""Generate tests/golden/reference_run.json by EXECUTING the reference's own scripts.  Run HERE (the build container,
where /root/reference exists); commit the output.  The GPU box never runs this.

    python tests/golden/make_golden_from_reference.py [/root/reference]

Each of /root/reference/{arbitrage,liquidation,two-asset}.py is run unmodified with `runpy.run_path`: its data literals,
its dense A_i matrices, its cvxpy expression graph and constraint list are the reference's own text.  Three modules the
image lacks are stood in for at import time:
  * `cvxpy`             -> oracle/cvxpy_shim.py (same modelling API subset, scipy SLSQP with exact derivatives instead of
                           ECOS/Clarabel; the program is convex, so the optimum is the solver-independent part)
  * `matplotlib.pyplot` -> a recorder that swallows the plotting calls of two-asset.py:102-118
  * `latexify`          -> the same (it only sets matplotlib rcParams, latexify.py:8-73)
What is stored is exactly what the scripts read back after `prob.solve()`: prob.value / psi.value / deltas[i].value /
lambdas[i].value (arbitrage.py:84, liquidation.py:87) and, per t, psi.value[2], lambdas[k].value - deltas[k].value and
obj.value (two-asset.py:93-100), plus the data literals the scripts define (so the tests can check that
cfmm_routing_code_b200.instances restates them exactly).
"""
import contextlib
import io
import json
import os
import runpy
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import cvxpy_shim                              # noqa: E402


class _Swallow(types.ModuleType):
    """module stand-in: any attribute is a function that accepts anything and returns None"""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *a, **k: None


def run_reference_script(path):
    """globals of the script after it ran to completion"""
    saved = {k: sys.modules.get(k) for k in ("cvxpy", "matplotlib", "matplotlib.pyplot", "latexify")}
    mpl = _Swallow("matplotlib")
    mpl.pyplot = _Swallow("matplotlib.pyplot")
    lat = _Swallow("latexify")
    sys.modules.update({"cvxpy": cvxpy_shim, "matplotlib": mpl, "matplotlib.pyplot": mpl.pyplot, "latexify": lat})
    try:
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            g = runpy.run_path(path, run_name="__main__")
        g["__stdout__"] = buf.getvalue()
        return g
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _lists(xs):
    return [[float(v) for v in np.asarray(x).reshape(-1)] for x in xs]


def _data(g):
    d = dict(n_tokens=len(g["global_indices"]), local_indices=[[int(t) for t in l] for l in g["local_indices"]],
             reserves=_lists(g["reserves"]), fees=[float(f) for f in g["fees"]])
    if "market_value" in g:
        d["market_value"] = [float(v) for v in g["market_value"]]
    if "current_assets" in g:
        d["current_assets"] = [float(v) for v in np.asarray(g["current_assets"]).reshape(-1)]
    return d


def main(ref="/root/reference"):
    out = {"how": "the reference's own scripts executed by runpy with oracle/cvxpy_shim.py standing in for cvxpy "
                  "(scipy SLSQP back end, exact derivatives); see tests/golden/make_golden_from_reference.py",
           "shim": cvxpy_shim.__version__}
    g = run_reference_script(os.path.join(ref, "arbitrage.py"))
    out["arbitrage"] = dict(data=_data(g), status=g["prob"].status, value=float(g["prob"].value),
                            psi=[float(v) for v in g["psi"].value], deltas=_lists(d.value for d in g["deltas"]),
                            lambdas=_lists(l.value for l in g["lambdas"]), stdout=g["__stdout__"].strip())
    g = run_reference_script(os.path.join(ref, "liquidation.py"))
    out["liquidation"] = dict(data=_data(g), status=g["prob"].status, value=float(g["psi"].value[4]),
                              prob_value=float(g["prob"].value), psi=[float(v) for v in g["psi"].value],
                              deltas=_lists(d.value for d in g["deltas"]), lambdas=_lists(l.value for l in g["lambdas"]),
                              stdout=g["__stdout__"].strip())
    g = run_reference_script(os.path.join(ref, "two-asset.py"))
    av = g["all_values"]            # all_values[k][:, j] = lambdas[k].value - deltas[k].value at amounts[j]  (two-asset.py:93-94)
    out["two_asset"] = dict(data=_data(g), amounts=[float(t) for t in g["amounts"]], u_t=[float(u) for u in g["u_t"]],
                            flows=[[[float(v) for v in av[k][:, j]] for k in range(len(av))] for j in range(len(g["amounts"]))],
                            stdout_tail=g["__stdout__"].strip().splitlines()[-6:])
    dst = os.path.join(HERE, "reference_run.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", dst)
    print("arbitrage.py  :", out["arbitrage"]["stdout"])
    print("liquidation.py:", out["liquidation"]["stdout"])
    print("two-asset.py  : u(0) =", out["two_asset"]["u_t"][0], " u(50) =", out["two_asset"]["u_t"][-1])


if __name__ == "__main__":
    main(*sys.argv[1:2])
