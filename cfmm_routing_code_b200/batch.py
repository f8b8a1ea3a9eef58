"""Batches of small routing problems: every ``prob.solve()`` of the batch in ONE kernel launch.

The reference solves its problems one at a time and, for the sweep of two-asset.py:40-100, rebuilds the cvxpy problem
for each of the 50 trade sizes.  Here the pools (the literals of two-asset.py:5-31) are uploaded once as CSR arrays and
``cfmm_batch_solve`` (csrc/cfmm_small.cu) runs one complete dual solve per GPU thread.  Nothing here has a CPU path.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .pools import HostPools, KIND_GEOMEAN_HOST
from .solver import default_nu0

NTOK_MAX = 64      # cfmm_small::NTOK_MAX
ARITY_MAX = 8      # cfmm_small::KMAX


def batch_applicable(hp: HostPools) -> bool:
    """Whether the per-thread solver covers this pool set (token count and arities of the reference's instances do)."""
    if hp.n_tokens > NTOK_MAX or hp.m == 0:
        return False
    return int(np.diff(hp.pool_ptr).max()) <= ARITY_MAX


class CsrStore:
    """The CSR pool arrays in HBM (what `local_indices`, `reserves`, `fees` of arbitrage.py:6-28 become)."""

    def __init__(self, hp: HostPools, device="cuda"):
        if not torch.cuda.is_available():
            raise _lib.CfmmError("cfmm_routing_code_b200 needs a CUDA device (there is no CPU fallback)")
        hp.validate()
        if not batch_applicable(hp):
            raise ValueError(f"the batched solver takes at most {NTOK_MAX} tokens and arity <= {ARITY_MAX}")
        self.lib = _lib.load()
        self.hp, self.device = hp, torch.device(device)
        self.n_tokens, self.m, self.nnz = hp.n_tokens, hp.m, int(len(hp.tok_idx))
        dev = self.device
        per_slot_kind = np.repeat(np.asarray(hp.kind), np.diff(hp.pool_ptr))
        w = np.where(per_slot_kind == KIND_GEOMEAN_HOST, hp.weights, 1.0)      # log(R/w) is a weighted-pool quantity
        self.pool_ptr = torch.as_tensor(np.ascontiguousarray(hp.pool_ptr, np.int64), device=dev)
        self.tok = torch.as_tensor(np.ascontiguousarray(hp.tok_idx, np.int32), device=dev)
        self.R = torch.as_tensor(np.ascontiguousarray(hp.reserves, np.float64), device=dev)
        self.w = torch.as_tensor(np.ascontiguousarray(hp.weights, np.float64), device=dev)
        self.logrw = torch.log(torch.clamp(self.R, min=1e-300) / torch.as_tensor(w, device=dev))
        self.gamma = torch.as_tensor(np.ascontiguousarray(hp.gamma, np.float64), device=dev)
        self.kind = torch.as_tensor(np.ascontiguousarray(hp.kind, np.uint8), device=dev)
        self.c_pools = _lib.CsrPools(self.n_tokens, self.m, self.nnz, self.pool_ptr.data_ptr(), self.tok.data_ptr(),
                                     self.R.data_ptr(), self.w.data_ptr(), self.logrw.data_ptr(),
                                     self.gamma.data_ptr(), self.kind.data_ptr())
        self._work = None

    def work(self, n_problems: int, nnz_max: int = 0) -> torch.Tensor:
        nbytes = self.lib.cfmm_batch_solve_work_bytes(C.byref(self.c_pools), n_problems, nnz_max)
        if nbytes < 0:
            _lib.check(int(nbytes), "cfmm_batch_solve_work_bytes")
        if self._work is None or self._work.numel() < nbytes:
            self._work = torch.empty(max(int(nbytes), 8), dtype=torch.uint8, device=self.device)
        return self._work


SMALL_BATCH = 512            # up to this many problems a warp per problem beats a thread per problem
LANES_SMALL_BATCH = 32


@torch.no_grad()
def solve_batch_device(store: CsrStore, c: torch.Tensor, a: torch.Tensor, flags: torch.Tensor, nu: torch.Tensor,
                       tol: float = 1e-8, want_trades: bool = True, pool_range: Optional[torch.Tensor] = None,
                       max_outer: int = 60, max_inner: int = 100, nnz_max: int = 0, lanes: Optional[int] = None):
    """Device-resident form: c, a [B, n] f64, flags [B, n] u8, nu [B, n] f64 (start prices, overwritten with the
    solution).  Returns (psi [B, n], stats [B, 8], delta, lambda [B, nnz] or None).  Asynchronous on the current stream.
    pool_range [B, 2] int64 (device): problem p uses pools [lo, hi) -- disjoint problems packed into one CSR array; then
    delta / lambda are [1, nnz] and nnz_max (slots of the largest problem) sizes the workspace."""
    B, n = c.shape
    if n != store.n_tokens:
        raise ValueError("utilities must have one entry per token")
    dev = store.device
    f64 = dict(dtype=torch.float64, device=dev)
    psi = torch.empty(B, n, **f64)
    stats = torch.empty(B, 8, **f64)
    shared = pool_range is None
    delta = lam = None
    if want_trades:
        delta = torch.zeros(B if shared else 1, store.nnz, **f64)
        lam = torch.zeros(B if shared else 1, store.nnz, **f64)
    if lanes is None:
        # a problem per warp for small batches (latency: the two-asset.py sweep of 50 problems takes 0.75 ms that way
        # against 1.14 ms with a thread per problem), a problem per thread for large ones (throughput: 4096 problems
        # 1.27 ms against 2.67 ms) -- measured on B200, profiles/r2a_batch_lanes.txt
        lanes = LANES_SMALL_BATCH if B <= SMALL_BATCH else 1
    _lib.check(store.lib.cfmm_set_batch_lanes(int(lanes)), "cfmm_set_batch_lanes")
    work = store.work(B, nnz_max)
    batch = _lib.Batch(B, None if shared else pool_range.data_ptr(), c.data_ptr(), a.data_ptr(), flags.data_ptr(),
                       nu.data_ptr(), psi.data_ptr(), stats.data_ptr(),
                       delta.data_ptr() if want_trades else None, lam.data_ptr() if want_trades else None,
                       store.nnz if shared else 0)
    prm = _lib.BatchParams(float(tol), 0.1, 1e-4, 0.5, 1e-12, int(max_outer), int(max_inner))
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(store.lib.cfmm_batch_solve(C.byref(store.c_pools), C.byref(batch), C.byref(prm), work.data_ptr(), st),
               "cfmm_batch_solve")
    return psi, stats, delta, lam


def pack_utilities(utilities: Sequence, n: int, nu0=None):
    """Host-side [B, n] arrays of the utilities' linear+box form (api.Arbitrage / Liquidate / Swap .spec())."""
    B = len(utilities)
    c = np.empty((B, n)); a = np.empty((B, n)); fl = np.empty((B, n), np.uint8); nu = np.empty((B, n))
    for p, u in enumerate(utilities):
        sp = u.spec(n)
        c[p] = sp.c; a[p] = sp.a
        fl[p] = np.asarray(sp.eq, np.uint8) | (np.asarray(sp.pinned, np.uint8) << 1)
        nu[p] = default_nu0(sp) if nu0 is None else np.asarray(nu0[p] if np.ndim(nu0) == 2 else nu0, float)
    return c, a, fl, nu


def solve_batch(hp: HostPools, utilities: Sequence, nu0=None, tol: float = 1e-8, device="cuda",
                want_trades: bool = True, store: Optional[CsrStore] = None, max_inner: int = 100):
    """All problems (same pools, one utility each) in one launch.  Returns a list of api.Result."""
    from .api import Result, infeasible_suspected
    t0 = time.perf_counter()
    from .api import _check_structurally_feasible
    for u in utilities:
        _check_structurally_feasible(hp, u.spec(hp.n_tokens))
    store = store or CsrStore(hp, device=device)
    n, dev = hp.n_tokens, store.device
    c, a, fl, nu = pack_utilities(utilities, n, nu0)
    up = lambda x: torch.as_tensor(x, device=dev)
    nu_d = up(nu)
    psi, stats, delta, lam = solve_batch_device(store, up(c), up(a), up(fl), nu_d, tol=tol, want_trades=want_trades,
                                                max_inner=max_inner)
    stats_h = stats.cpu().numpy()              # the one synchronisation of the call
    psi_h, nu_h = psi.cpu().numpy(), nu_d.cpu().numpy()
    if want_trades:
        d_h, l_h = delta.cpu().numpy(), lam.cpu().numpy()
    wall = time.perf_counter() - t0
    ptr = hp.pool_ptr
    names = {0: "optimal", 1: "max_iter", 2: "stalled", 3: "rejected"}
    out: List = []
    for p in range(len(utilities)):
        s = stats_h[p]
        deltas = [d_h[p, ptr[i]:ptr[i + 1]] for i in range(hp.m)] if want_trades else []
        lambdas = [l_h[p, ptr[i]:ptr[i + 1]] for i in range(hp.m)] if want_trades else []
        status = names[int(s[7])]
        if status in ("max_iter", "stalled") and infeasible_suspected(utilities[p].spec(hp.n_tokens), nu_h[p], psi_h[p], status):
            status = "infeasible"
        out.append(Result(value=float(s[0]), psi=psi_h[p], deltas=deltas, lambdas=lambdas, nu=nu_h[p],
                          dual_value=float(s[1]), gap=float(s[2]), primal_infeas=float(s[3]), iters=int(s[5]),
                          evals=int(s[6]), hvps=0, status=status, wall_s=wall, info=None))
    return out


def pack_problems(problems: Sequence):
    """Concatenate independent small problems [(HostPools, utility), ...] into one CSR array + per-problem pool ranges.
    Token indices stay local to each problem; problems with fewer tokens than the widest are padded with tokens no pool
    touches, pinned at price 1 (objective-only, zero coefficient)."""
    n = max(hp.n_tokens for hp, _ in problems)
    B = len(problems)
    ptr = [np.zeros(1, np.int64)]
    ranges = np.empty((B, 2), np.int64)
    c = np.zeros((B, n)); a = np.zeros((B, n)); fl = np.full((B, n), 2, np.uint8); nu = np.ones((B, n))
    m0, off0, nnz_max = 0, 0, 0
    for p, (hp, u) in enumerate(problems):
        hp.validate()
        ptr.append(np.asarray(hp.pool_ptr[1:], np.int64) + off0)
        ranges[p] = (m0, m0 + hp.m)
        m0 += hp.m
        off0 += int(hp.pool_ptr[-1])
        nnz_max = max(nnz_max, int(hp.pool_ptr[-1]))
        sp = u.spec(hp.n_tokens)
        k = hp.n_tokens
        c[p, :k] = sp.c; a[p, :k] = sp.a
        fl[p, :k] = np.asarray(sp.eq, np.uint8) | (np.asarray(sp.pinned, np.uint8) << 1)
        nu[p, :k] = default_nu0(sp)
        c[p, k:] = 1.0
    cat = lambda name, dt: np.concatenate([np.asarray(getattr(hp, name), dt) for hp, _ in problems])
    merged = HostPools(n, np.concatenate(ptr), cat("tok_idx", np.int32), cat("reserves", np.float64),
                       cat("weights", np.float64), cat("gamma", np.float64), cat("kind", np.uint8))
    return merged, ranges, c, a, fl, nu, nnz_max


def solve_many(problems: Sequence, tol: float = 1e-8, device="cuda", want_trades: bool = True):
    """Independent small problems -- each its own pools and utility, e.g. one per market or per block -- in ONE launch.
    problems: [(HostPools, utility), ...].  Returns a list of api.Result, in order."""
    from .api import Result
    t0 = time.perf_counter()
    if len(problems) == 0:
        return []
    merged, ranges, c, a, fl, nu, nnz_max = pack_problems(problems)
    store = CsrStore(merged, device=device)
    dev = store.device
    up = lambda x: torch.as_tensor(x, device=dev)
    nu_d = up(nu)
    psi, stats, delta, lam = solve_batch_device(store, up(c), up(a), up(fl), nu_d, tol=tol, want_trades=want_trades,
                                                pool_range=up(ranges), nnz_max=nnz_max)
    stats_h, psi_h, nu_h = stats.cpu().numpy(), psi.cpu().numpy(), nu_d.cpu().numpy()
    if want_trades:
        d_h, l_h = delta.cpu().numpy()[0], lam.cpu().numpy()[0]
    wall = time.perf_counter() - t0
    names = {0: "optimal", 1: "max_iter", 2: "stalled", 3: "rejected"}
    out: List = []
    ptr = merged.pool_ptr
    for p, (hp, _) in enumerate(problems):
        s, k = stats_h[p], hp.n_tokens
        lo, hi = ranges[p]
        deltas = [d_h[ptr[i]:ptr[i + 1]] for i in range(lo, hi)] if want_trades else []
        lambdas = [l_h[ptr[i]:ptr[i + 1]] for i in range(lo, hi)] if want_trades else []
        out.append(Result(value=float(s[0]), psi=psi_h[p, :k], deltas=deltas, lambdas=lambdas, nu=nu_h[p, :k],
                          dual_value=float(s[1]), gap=float(s[2]), primal_infeas=float(s[3]), iters=int(s[5]),
                          evals=int(s[6]), hvps=0, status=names[int(s[7])], wall_s=wall, info=None))
    return out
