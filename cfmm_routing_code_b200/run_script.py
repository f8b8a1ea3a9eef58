"""Run a cvxpy script of the reference's kind UNMODIFIED on the B200 path.

    python -m cfmm_routing_code_b200.run_script /path/to/arbitrage.py [--plots]

`import cvxpy as cp` inside the script resolves to `cfmm_routing_code_b200.cvxpy_compat`, so its `prob.solve()`
(arbitrage.py:81-82, liquidation.py:84-85, two-asset.py:90-91) is the CUDA solve and everything the script prints or
reads afterwards comes from it.  Without --plots the script's plotting imports (`matplotlib.pyplot`, `latexify`:
two-asset.py:1-5, :102-118) are replaced by no-ops, so a headless GPU box needs neither.
"""
from __future__ import annotations

import runpy
import sys
import types

from . import cvxpy_compat


class _NoOp(types.ModuleType):
    """a module whose every attribute is a function that accepts anything and returns None"""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *a, **k: None


def run(path: str, plots: bool = False, run_name: str = "__main__") -> dict:
    """Execute the script at `path`; returns its globals (prob, psi, deltas, lambdas, ... as the script named them)."""
    names = ["cvxpy"] + ([] if plots else ["matplotlib", "matplotlib.pyplot", "latexify"])
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules["cvxpy"] = cvxpy_compat
    if not plots:
        mpl = _NoOp("matplotlib")
        mpl.pyplot = _NoOp("matplotlib.pyplot")
        sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": mpl.pyplot, "latexify": _NoOp("latexify")})
    try:
        return runpy.run_path(path, run_name=run_name)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    plots = "--plots" in argv
    paths = [a for a in argv if not a.startswith("--")]
    if len(paths) != 1:
        sys.stderr.write(__doc__)
        return 2
    run(paths[0], plots=plots)
    return 0


if __name__ == "__main__":
    sys.exit(main())
