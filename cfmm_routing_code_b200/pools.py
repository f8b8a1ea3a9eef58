"""Device-resident pool storage and the per-evaluation kernel calls.

Replaces the reference's dense local->global matrices A_i (arbitrage.py:42-48) with index rows, and
its python lists ``reserves`` / ``fees`` (arbitrage.py:14-28) with slot-major SoA buckets in HBM,
one bucket per (kind, arity).  All compute goes through libcfmm_b200.so; nothing here has a CPU path.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib

KIND_GEOMEAN_HOST = 0   # host CSR convention: 0 = (weighted) geometric mean, 1 = constant sum
KIND_SUM_HOST = 1


@dataclasses.dataclass
class HostPools:
    """CSR problem data on the host (numpy)."""
    n_tokens: int
    pool_ptr: np.ndarray   # int64 [m+1]
    tok_idx: np.ndarray    # int32 [nnz]
    reserves: np.ndarray   # f64 [nnz]
    weights: np.ndarray    # f64 [nnz]  normalised per pool; ignored for constant-sum
    gamma: np.ndarray      # f64 [m]
    kind: np.ndarray       # uint8 [m]

    @property
    def m(self) -> int:
        return int(len(self.gamma))

    @staticmethod
    def from_lists(n_tokens, local_indices, reserves, fees, kinds, weights=None) -> "HostPools":
        """From the reference's literals: local_indices / reserves / fees (arbitrage.py:6-28) plus
        which cvxpy atom constrains each pool (arbitrage.py:63-74): 'geomean' (with weights[i] = the
        ``p=`` vector), 'product' (cp.geo_mean on 2 tokens) or 'sum'."""
        m = len(local_indices)
        if not (len(reserves) == len(fees) == len(kinds) == m):
            raise ValueError("local_indices, reserves, fees, kinds must have one entry per pool")
        ptr = [0]; idx: List[int] = []; res: List[float] = []; wts: List[float] = []; kd: List[int] = []
        for i, l in enumerate(local_indices):
            k = len(l)
            if len(reserves[i]) != k:
                raise ValueError(f"pool {i}: {len(reserves[i])} reserves for {k} tokens")
            if len(set(int(t) for t in l)) != k:
                raise ValueError(f"pool {i}: repeated token in local_indices")
            ptr.append(ptr[-1] + k)
            idx += [int(t) for t in l]
            res += [float(x) for x in reserves[i]]
            if kinds[i] == "sum":
                if k != 2:
                    raise ValueError("constant-sum pools must have 2 tokens (as arbitrage.py:11)")
                kd.append(KIND_SUM_HOST); wts += [0.0] * k
            elif kinds[i] in ("geomean", "product"):
                w = np.ones(k) if (weights is None or weights[i] is None) else np.asarray(weights[i], float)
                if len(w) != k or np.any(w <= 0):
                    raise ValueError(f"pool {i}: bad weights")
                kd.append(KIND_GEOMEAN_HOST); wts += list(w / w.sum())
            else:
                raise ValueError(f"pool {i}: unknown kind {kinds[i]!r}")
        return HostPools(int(n_tokens), np.asarray(ptr, np.int64), np.asarray(idx, np.int32),
                         np.asarray(res, np.float64), np.asarray(wts, np.float64),
                         np.asarray(fees, np.float64), np.asarray(kd, np.uint8))

    @staticmethod
    def from_pairs(n_tokens, idx, reserves, gamma) -> "HostPools":
        """m constant-product pools given as (m,2) arrays."""
        m = len(gamma)
        return HostPools(int(n_tokens), np.arange(0, 2 * m + 1, 2, dtype=np.int64),
                         np.ascontiguousarray(idx, np.int32).reshape(-1),
                         np.ascontiguousarray(reserves, np.float64).reshape(-1),
                         np.full(2 * m, 0.5), np.ascontiguousarray(gamma, np.float64), np.zeros(m, np.uint8))

    def validate(self):
        if np.any(self.reserves <= 0) or not np.all(np.isfinite(self.reserves)):
            raise ValueError("reserves must be positive and finite")
        if np.any(self.gamma <= 0) or np.any(self.gamma > 1):
            raise ValueError("fees (gamma) must lie in (0, 1]")
        if self.tok_idx.min(initial=0) < 0 or self.tok_idx.max(initial=0) >= self.n_tokens:
            raise ValueError("token index out of range")


def split_buckets(hp: HostPools, rank: int = 0, world: int = 1):
    """Group pools by (kind, arity); with world>1 keep this rank's contiguous block of each group."""
    ar = np.diff(hp.pool_ptr)
    out = []
    w2 = hp.weights[hp.pool_ptr[:-1]] if hp.m else np.zeros(0)
    is_cp = (hp.kind == KIND_GEOMEAN_HOST) & (ar == 2)
    if hp.m:
        is_cp &= (w2 == 0.5) & (hp.weights[np.minimum(hp.pool_ptr[:-1] + 1, len(hp.weights) - 1)] == 0.5)
    keys = []
    if is_cp.any():
        keys.append((_lib.KIND_PRODUCT, 2, np.nonzero(is_cp)[0]))
    cs = hp.kind == KIND_SUM_HOST
    if cs.any():
        if np.any(ar[cs] != 2):
            raise ValueError("constant-sum pools must have 2 tokens")
        keys.append((_lib.KIND_SUM, 2, np.nonzero(cs)[0]))
    gm = (hp.kind == KIND_GEOMEAN_HOST) & ~is_cp
    for k in sorted(set(ar[gm].tolist())):
        if k < 2 or k > 32:
            raise ValueError(f"weighted pools support 2..32 tokens, got {k}")
        keys.append((_lib.KIND_GEOMEAN, int(k), np.nonzero(gm & (ar == k))[0]))
    for kind, k, sel in keys:
        if world > 1:
            lo = (len(sel) * rank) // world
            hi = (len(sel) * (rank + 1)) // world
            sel = sel[lo:hi]
        off = hp.pool_ptr[sel][None, :] + np.arange(k)[:, None]        # (k, m_b) slot-major
        out.append(dict(kind=kind, arity=k, sel=sel, off=off))
    return out


TILE = 1024     # pools per TMA tile (csrc kTile): slot stride is padded to a multiple of it


def _padded(arr2d: np.ndarray, stride: int, fill) -> np.ndarray:
    k, m = arr2d.shape
    out = np.full((k, stride), fill, dtype=arr2d.dtype)
    out[:, :m] = arr2d
    return out


class DeviceBucket:
    def __init__(self, hp: HostPools, spec, device):
        self.kind = spec["kind"]; self.arity = spec["arity"]
        self.sel = spec["sel"]; self.off = spec["off"]
        self.m = int(len(self.sel))
        self.stride = max(TILE, -(-self.m // TILE) * TILE)
        f64 = dict(dtype=torch.float64, device=device)
        R = hp.reserves[self.off]
        self.reserves = torch.as_tensor(_padded(R, self.stride, 1.0), **f64)
        self.tok_idx = torch.as_tensor(_padded(hp.tok_idx[self.off].astype(np.int32), self.stride, 0),
                                       dtype=torch.int32, device=device)
        self.gamma = torch.as_tensor(_padded(hp.gamma[self.sel][None, :], self.stride, 1.0)[0], **f64)
        self.weights = self.logrw = self.theta_bar = None
        if self.kind == _lib.KIND_GEOMEAN:
            W = hp.weights[self.off]
            self.weights = torch.as_tensor(_padded(W, self.stride, 1.0), **f64)
            self.logrw = torch.as_tensor(_padded(np.log(R / W), self.stride, 0.0), **f64)
        if self.kind == _lib.KIND_SUM:
            self.theta_bar = torch.zeros((2, self.stride), **f64)
        self.delta = self.lam = self.hcoef = self.hmask = None
        self._device = device
        self.c_bucket = _lib.Bucket(
            self.kind, self.arity, self.m, self.stride, self.reserves.data_ptr(), self.tok_idx.data_ptr(),
            self.gamma.data_ptr(),
            self.weights.data_ptr() if self.weights is not None else None,
            self.logrw.data_ptr() if self.logrw is not None else None,
            self.theta_bar.data_ptr() if self.theta_bar is not None else None)

    def bytes_resident(self) -> int:
        n = 0
        for t in (self.reserves, self.tok_idx, self.gamma, self.weights, self.logrw, self.theta_bar):
            if t is not None:
                n += t.numel() * t.element_size()
        return n

    def out_struct(self, trades: bool, hess: bool):
        f64 = dict(dtype=torch.float64, device=self._device)
        if trades and self.delta is None:
            self.delta = torch.zeros((self.arity, self.stride), **f64)
            self.lam = torch.zeros((self.arity, self.stride), **f64)
        if hess and self.hcoef is None:
            self.hcoef = torch.zeros(self.stride, **f64)
            self.hmask = torch.zeros(self.stride, dtype=torch.int32, device=self._device)
        return _lib.EvalOut(self.delta.data_ptr() if trades else None, self.lam.data_ptr() if trades else None,
                            self.hcoef.data_ptr() if hess else None, self.hmask.data_ptr() if hess else None)


class PoolStore:
    """All pools of one problem (or one rank's shard of them), resident on one GPU."""

    def __init__(self, hp: HostPools, device="cuda", rank: int = 0, world: int = 1, validate: bool = True):
        if validate:
            hp.validate()
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CfmmError("PoolStore needs a CUDA device: there is no CPU path in this package")
        self.n_tokens = int(hp.n_tokens)
        self.m_total = hp.m
        self.pool_ptr = hp.pool_ptr
        self.rank, self.world = rank, world
        self.buckets = [DeviceBucket(hp, s, self.device) for s in split_buckets(hp, rank, world)]
        self.m_local = sum(b.m for b in self.buckets)
        self.has_sum = bool(np.any(hp.kind == KIND_SUM_HOST))
        self.has_geomean = any(b.kind == _lib.KIND_GEOMEAN for b in self.buckets)
        f64 = dict(dtype=torch.float64, device=self.device)
        self._acc = torch.zeros(self.n_tokens + 1, **f64)      # [psi | arb], the one all-reduced buffer
        self._y = torch.zeros(self.n_tokens, **f64)
        self._move = torch.zeros(1, **f64)
        self.evals = 0
        self.hvps = 0

    # -- helpers -------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def bytes_resident(self) -> int:
        return sum(b.bytes_resident() for b in self.buckets)

    def algorithmic_bytes_per_eval(self) -> int:
        """32 B per 2-token pool, 28k+12 per weighted pool, + nu, psi, arb."""
        n = 0
        for b in self.buckets:
            n += b.m * (32 if b.arity == 2 and b.kind != _lib.KIND_GEOMEAN else 28 * b.arity + 12)
        return n + 16 * self.n_tokens + 8

    # -- the hot path --------------------------------------------------------------------------
    def evaluate(self, nu: torch.Tensor, eps: float = 0.0, trades: bool = False, hess: bool = False):
        """psi(nu) (n_tokens) and arb(nu) (1) for this rank's pools, as views into one (n+1) buffer."""
        st = self._stream()
        acc = self._acc
        _lib.check(self.lib.cfmm_zero(acc.data_ptr(), acc.numel() * 8, st), "cfmm_zero")
        lognu = torch.log(nu) if self.has_geomean else None
        for b in self.buckets:
            out = b.out_struct(trades, hess) if (trades or hess) else None
            rc = self.lib.cfmm_arb_eval(C.byref(b.c_bucket), self.n_tokens, nu.data_ptr(),
                                        lognu.data_ptr() if lognu is not None else None, float(eps),
                                        acc.data_ptr(), acc.data_ptr() + 8 * self.n_tokens,
                                        C.byref(out) if out is not None else None, st)
            _lib.check(rc, "cfmm_arb_eval")
        self.evals += 1
        return acc

    def hvp(self, vt: torch.Tensor) -> torch.Tensor:
        st = self._stream()
        y = self._y
        _lib.check(self.lib.cfmm_zero(y.data_ptr(), y.numel() * 8, st), "cfmm_zero")
        for b in self.buckets:
            rc = self.lib.cfmm_hvp(C.byref(b.c_bucket), self.n_tokens, b.hcoef.data_ptr(),
                                   b.hmask.data_ptr(), vt.data_ptr(), y.data_ptr(), st)
            _lib.check(rc, "cfmm_hvp")
        self.hvps += 1
        return y

    def hess_diag(self) -> torch.Tensor:
        st = self._stream()
        d = torch.zeros(self.n_tokens, dtype=torch.float64, device=self.device)
        for b in self.buckets:
            _lib.check(self.lib.cfmm_hess_diag(C.byref(b.c_bucket), self.n_tokens, b.hcoef.data_ptr(),
                                               b.hmask.data_ptr(), d.data_ptr(), st), "cfmm_hess_diag")
        return d

    def hess_dense(self) -> torch.Tensor:
        st = self._stream()
        H = torch.zeros((self.n_tokens, self.n_tokens), dtype=torch.float64, device=self.device)
        for b in self.buckets:
            _lib.check(self.lib.cfmm_hess_dense(C.byref(b.c_bucket), self.n_tokens, b.hcoef.data_ptr(),
                                                b.hmask.data_ptr(), H.data_ptr(), st), "cfmm_hess_dense")
        return H

    def update_multipliers(self) -> torch.Tensor:
        """theta_bar <- fills of the last trades=True evaluation; returns max relative change (device)."""
        st = self._stream()
        self._move.zero_()
        for b in self.buckets:
            if b.kind == _lib.KIND_SUM:
                _lib.check(self.lib.cfmm_sum_update_multipliers(C.byref(b.c_bucket), b.lam.data_ptr(),
                                                                b.theta_bar.data_ptr(), self._move.data_ptr(), st),
                           "cfmm_sum_update_multipliers")
        return self._move

    def reset_multipliers(self):
        for b in self.buckets:
            if b.theta_bar is not None:
                b.theta_bar.zero_()

    def gather_trades(self):
        """Delta, Lambda of the last trades=True evaluation, CSR order of the ORIGINAL pools (host).
        Entries of pools living on other ranks are left at zero."""
        nnz = int(self.pool_ptr[-1])
        delta = np.zeros(nnz); lam = np.zeros(nnz)
        for b in self.buckets:
            if b.m == 0:
                continue
            delta[b.off.ravel()] = b.delta[:, :b.m].cpu().numpy().ravel()
            lam[b.off.ravel()] = b.lam[:, :b.m].cpu().numpy().ravel()
        return delta, lam
