"""TEST-ONLY evaluator: the solver's evaluator protocol implemented with the numpy oracle on CPU tensors.

Lets the `-m "not gpu"` suite exercise the product's host logic (solver.py, sharding, all-reduce plumbing)
without a GPU.  The product never imports this."""
import numpy as np
import torch

from oracle import cfmm_oracle as O


def oracle_pools(hp):
    return O.Pools(hp.n_tokens, hp.pool_ptr, hp.tok_idx, hp.reserves, hp.weights, hp.gamma, hp.kind)


class OracleEvaluator:
    def __init__(self, hp, rank=0, world=1):
        P = oracle_pools(hp)
        if world > 1:      # contiguous shard of the pool list
            lo, hi = (P.m * rank) // world, (P.m * (rank + 1)) // world
            ptr = P.pool_ptr
            s = slice(ptr[lo], ptr[hi])
            P = O.Pools(P.n_tokens, ptr[lo:hi + 1] - ptr[lo], P.tok_idx[s], P.reserves[s], P.weights[s],
                        P.gamma[lo:hi], P.kind[lo:hi])
        self.P = P
        self.bk = O.Buckets(P)
        self.n_tokens = P.n_tokens
        self.has_sum = bool(np.any(hp.kind == O.KIND_CONST_SUM))
        self.device = torch.device("cpu")
        self.evals = 0
        self.hvps = 0
        self._hs = None
        self._last = None
        for g in self.bk.groups:
            if g["kind"] == O.KIND_CONST_SUM:
                g["theta_bar"] = np.zeros_like(g["R"])

    def evaluate(self, nu, eps=0.0, trades=False, hess=False):
        ev = O.evaluate(self.bk, nu.numpy(), eps, want_trades=trades, want_hess=hess)
        self.evals += 1
        if hess:
            self._hs = ev["hess_scaled"]
        if trades:
            self._last = ev
        return torch.as_tensor(np.concatenate([ev["psi"], [ev["arb"]]]))

    def hvp(self, vt):
        self.hvps += 1
        return torch.as_tensor(self._hs @ vt.numpy())

    def hess_diag(self):
        return torch.as_tensor(np.diag(self._hs).copy())

    def hess_dense(self):
        return torch.as_tensor(self._hs.copy())

    def reset_multipliers(self):
        for g in self.bk.groups:
            if g["kind"] == O.KIND_CONST_SUM:
                g["theta_bar"][:] = 0.0

    def update_multipliers(self):
        move = 0.0
        for g in self.bk.groups:
            if g["kind"] == O.KIND_CONST_SUM and len(g["sel"]):
                th = self._last["lam"][g["off"]]
                move = max(move, float(np.max(np.abs(th - g["theta_bar"]) / g["R"])))
                g["theta_bar"] = th.copy()
        return torch.tensor([move], dtype=torch.float64)

    def gather_trades(self):
        return self._last["delta"], self._last["lam"]
