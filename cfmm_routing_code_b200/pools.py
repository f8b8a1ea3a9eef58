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
KIND_BOUNDED_HOST = 3   # constant product on virtual reserves (reserves + offsets), real reserves >= 0; offsets ride in `weights`


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
            elif kinds[i] == "bounded_product":
                # not a reference atom: one Uniswap-v3 tick range; weights[i] = the two virtual-reserve offsets
                o = None if (weights is None or weights[i] is None) else np.asarray(weights[i], float)
                if k != 2 or o is None or len(o) != 2 or np.any(o < 0) or not np.all(np.isfinite(o)):
                    raise ValueError(f"pool {i}: bounded_product needs 2 tokens and 2 non-negative offsets in weights[i]")
                kd.append(KIND_BOUNDED_HOST); wts += list(o)
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
        hp = HostPools(int(n_tokens), np.arange(0, 2 * m + 1, 2, dtype=np.int64),
                       np.ascontiguousarray(idx, np.int32).reshape(-1),
                       np.ascontiguousarray(reserves, np.float64).reshape(-1),
                       np.full(2 * m, 0.5), np.ascontiguousarray(gamma, np.float64), np.zeros(m, np.uint8))
        hp._uniform_product = True          # known structure: split_buckets / validation take the device-side path
        return hp

    def pin_memory(self) -> "HostPools":
        """Move the arrays into page-locked host memory (in place) so PoolStore uploads run at full PCIe speed."""
        keep = []
        for name in ("tok_idx", "reserves", "weights", "gamma", "kind", "pool_ptr"):
            t = torch.from_numpy(np.ascontiguousarray(getattr(self, name))).pin_memory()
            keep.append(t)
            setattr(self, name, t.numpy())
        self._pinned = keep                  # the numpy views borrow these tensors' storage
        return self

    def validate(self):
        slot_kind = np.repeat(np.asarray(self.kind), np.diff(self.pool_ptr)) if np.any(self.kind == KIND_BOUNDED_HOST) else None
        virt = self.reserves if slot_kind is None else self.reserves + np.where(slot_kind == KIND_BOUNDED_HOST, self.weights, 0.0)
        lo_ok = np.all(self.reserves > 0) if slot_kind is None else np.all(self.reserves >= 0) and np.all(virt > 0)
        if not lo_ok or not np.all(np.isfinite(self.reserves)):
            raise ValueError("reserves must be positive and finite (bounded_product: >= 0 with positive virtual reserves)")
        if np.any(self.gamma <= 0) or np.any(self.gamma > 1):
            raise ValueError("fees (gamma) must lie in (0, 1]")
        if self.tok_idx.min(initial=0) < 0 or self.tok_idx.max(initial=0) >= self.n_tokens:
            raise ValueError("token index out of range")


class BucketSpec:
    """Which pools of a HostPools form one bucket.  `sel` = global pool indices (None = all pools, in order,
    the fast path for constant-product-only problems: no host-side gathers at all)."""

    def __init__(self, hp: HostPools, kind: int, arity: int, sel: Optional[np.ndarray], lo: int = 0,
                 hi: Optional[int] = None):
        self.hp, self.kind, self.arity, self._sel = hp, kind, arity, sel
        # sel is None: the contiguous block [lo, hi) of the pool list in its own order (a rank's shard, or everything)
        self.lo, self.hi = int(lo), int(hp.m if hi is None else hi)
        self._contiguous = sel is None
        self.m = (self.hi - self.lo) if sel is None else int(len(sel))
        self._off = None

    @property
    def identity(self) -> bool:
        """pools [lo, hi) in order: the raw host arrays are uploaded as they are and reordered on the GPU"""
        return self._contiguous

    @property
    def sel(self) -> np.ndarray:
        if self._sel is None:
            self._sel = np.arange(self.lo, self.hi, dtype=np.int64)
        return self._sel

    @property
    def off(self) -> np.ndarray:
        """(arity, m) CSR offsets of every slot of every pool of the bucket"""
        if self._off is None:
            self._off = self.hp.pool_ptr[self.sel][None, :] + np.arange(self.arity)[:, None]
        return self._off

    def subset(self, local_idx: np.ndarray) -> "BucketSpec":
        return BucketSpec(self.hp, self.kind, self.arity, self.sel[local_idx])


def split_buckets(hp: HostPools, rank: int = 0, world: int = 1) -> List["BucketSpec"]:
    """Group pools by (kind, arity); with world>1 keep this rank's contiguous block of each group."""
    m = hp.m
    if m == 0:
        return []
    if getattr(hp, "_uniform_product", False):
        lo, hi = (m * rank) // world, (m * (rank + 1)) // world
        return [BucketSpec(hp, _lib.KIND_PRODUCT, 2, None, lo, hi)]
    uniform_pairs = int(hp.pool_ptr[-1]) == 2 * m and (hp.pool_ptr[1] - hp.pool_ptr[0]) == 2 and \
        bool(np.all(np.diff(hp.pool_ptr[::max(1, m // 64)]) == 2 * max(1, m // 64))) and \
        bool(np.array_equal(hp.pool_ptr[:3], np.arange(0, 2 * min(m, 2) + 1, 2)[:3]))
    if uniform_pairs:
        uniform_pairs = bool(np.all(np.diff(hp.pool_ptr) == 2))
    if uniform_pairs and not hp.kind.any() and bool(np.all(hp.weights == 0.5)):
        lo, hi = (m * rank) // world, (m * (rank + 1)) // world
        return [BucketSpec(hp, _lib.KIND_PRODUCT, 2, None, lo, hi)]
    ar = np.diff(hp.pool_ptr)
    first = hp.pool_ptr[:-1]
    is_cp = (hp.kind == KIND_GEOMEAN_HOST) & (ar == 2)
    is_cp &= (hp.weights[first] == 0.5) & (hp.weights[np.minimum(first + 1, len(hp.weights) - 1)] == 0.5)
    keys = []
    if is_cp.any():
        keys.append((_lib.KIND_PRODUCT, 2, np.nonzero(is_cp)[0]))
    cs = hp.kind == KIND_SUM_HOST
    if cs.any():
        if np.any(ar[cs] != 2):
            raise ValueError("constant-sum pools must have 2 tokens")
        keys.append((_lib.KIND_SUM, 2, np.nonzero(cs)[0]))
    bp = hp.kind == KIND_BOUNDED_HOST
    if bp.any():
        if np.any(ar[bp] != 2):
            raise ValueError("bounded_product pools must have 2 tokens")
        keys.append((_lib.KIND_BOUNDED, 2, np.nonzero(bp)[0]))
    gm = (hp.kind == KIND_GEOMEAN_HOST) & ~is_cp
    for k in np.unique(ar[gm]).tolist():
        if k < 2 or k > 32:
            raise ValueError(f"weighted pools support 2..32 tokens, got {k}")
        keys.append((_lib.KIND_GEOMEAN, int(k), np.nonzero(gm & (ar == k))[0]))
    out = []
    for kind, k, sel in keys:
        if world > 1:
            lo = (len(sel) * rank) // world
            hi = (len(sel) * (rank + 1)) // world
            sel = sel[lo:hi]
        out.append(BucketSpec(hp, kind, k, sel))
    return out


TILE = 1024     # pools per TMA tile of the PLAIN buckets (csrc/cfmm_kernels.cu kTile): slot stride is padded to a multiple of it


def _padded(arr2d: np.ndarray, stride: int, fill) -> np.ndarray:
    k, m = arr2d.shape
    out = np.full((k, stride), fill, dtype=arr2d.dtype)
    out[:, :m] = arr2d
    return out


class DeviceBucket:
    def __init__(self, hp: HostPools, spec, device):
        self.kind = spec.kind; self.arity = spec.arity
        self.spec = spec
        self.m = spec.m
        self.stride = max(TILE, -(-self.m // TILE) * TILE)
        f64 = dict(dtype=torch.float64, device=device)
        self.weights = self.logrw = self.theta_bar = None
        if spec.identity and self.arity == 2 and int(hp.pool_ptr[-1]) == 2 * hp.m:   # uniform pairs in order: transpose on the GPU
            m, lo, hi = self.m, spec.lo, spec.hi
            self.reserves = torch.ones((2, self.stride), **f64)
            self.reserves[:, :m] = torch.from_numpy(hp.reserves[2 * lo:2 * hi]).to(device).view(m, 2).t()
            self.tok_idx = torch.zeros((2, self.stride), dtype=torch.int32, device=device)
            self.tok_idx[:, :m] = torch.from_numpy(hp.tok_idx[2 * lo:2 * hi]).to(device).view(m, 2).t()
            self.gamma = torch.ones(self.stride, **f64)
            self.gamma[:m] = torch.from_numpy(hp.gamma[lo:hi]).to(device)
            R = W = None
        else:
            R = hp.reserves[spec.off]
            self.reserves = torch.as_tensor(_padded(R, self.stride, 1.0), **f64)
            self.tok_idx = torch.as_tensor(_padded(hp.tok_idx[spec.off].astype(np.int32), self.stride, 0),
                                           dtype=torch.int32, device=device)
            self.gamma = torch.as_tensor(_padded(hp.gamma[spec.sel][None, :], self.stride, 1.0)[0], **f64)
        if self.kind == _lib.KIND_GEOMEAN:
            W = hp.weights[spec.off]
            self.weights = torch.as_tensor(_padded(W, self.stride, 1.0), **f64)
            self.logrw = torch.as_tensor(_padded(np.log(R / W), self.stride, 0.0), **f64)
        if self.kind == _lib.KIND_BOUNDED:                 # the virtual-reserve offsets ride in the weights slot
            self.weights = torch.as_tensor(_padded(hp.weights[spec.off], self.stride, 1.0), **f64)
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

    @property
    def sel(self) -> np.ndarray:
        return self.spec.sel

    @property
    def off(self) -> np.ndarray:
        return self.spec.off

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


def blocked_layout_info(lib):
    v = [C.c_int32() for _ in range(4)]
    _lib.check(lib.cfmm_blocked_layout_info(*[C.byref(x) for x in v]), "cfmm_blocked_layout_info")
    return tuple(int(x.value) for x in v)      # pools_per_tile, rows_stride, tok_stride, row_cap


def build_blocked_pairs(idx: torch.Tensor, n_tokens: int, P: int, rows_stride: int, tok_stride: int, row_cap: int):
    """Layout builder for cfmm_blocked_pairs (see csrc/cfmm_blocked.cuh).  idx: (2, m) int64 token ids on the
    device.  Pools are sorted by (token block of slot 0, token block of slot 1) and cut into tiles of P; each
    tile gets its distinct-token list, 16-bit local ids, and a CSR of rows (token, <= row_cap entries).
    Returns (order, residual, tables): `order` = bucket-local pool index at each blocked position, `residual` =
    pools left out because their tile would touch more than tok_stride tokens (they go to a plain bucket)."""
    dev = idx.device
    m = idx.shape[1]
    i64 = dict(dtype=torch.int64, device=dev)
    nb = max(1, int(round((m / P) ** 0.5)))
    a, b = idx[0], idx[1]
    # primary: (token block of slot 0, token block of slot 1); secondary: slot-0 token, so that the lanes of a warp
    # read the same nu_local entry (shared-memory broadcast) and scatter slot-0 flows to consecutive positions
    key = ((a * nb // n_tokens) * nb + (b * nb // n_tokens)) * n_tokens + a
    order = torch.argsort(key, stable=True)
    residual = []
    for _pass in range(4):
        mm = order.numel()
        if mm == 0:
            break
        pos = torch.arange(mm, **i64)
        ntiles = -(-mm // P)
        tile = pos // P
        he_tok = torch.cat([a[order], b[order]])
        he_tile = torch.cat([tile, tile])
        ck, perm = torch.sort(he_tile * n_tokens + he_tok, stable=True)
        uniq, inv, counts = torch.unique_consecutive(ck, return_inverse=True, return_counts=True)
        u_tile = uniq // n_tokens
        ntok = torch.bincount(u_tile, minlength=ntiles)
        bad = ntok > tok_stride
        if not bool(bad.any()):
            break
        if _pass == 3:                       # give up blocking: everything left goes to the plain bucket
            residual.append(order); order = order[:0]; mm = 0
            break
        keep = ~bad[tile]
        residual.append(order[~keep])
        order = order[keep]
    residual = torch.cat(residual) if residual else order[:0]
    if order.numel() == 0:
        return order, residual, None
    U = uniq.numel()
    tok_off = torch.cumsum(ntok, 0) - ntok
    ltok_u = torch.arange(U, **i64) - tok_off[u_tile]
    he_ltok = torch.empty(2 * mm, **i64)
    he_ltok[perm] = ltok_u[inv]
    M = ntiles * P
    lid = torch.zeros(M, dtype=torch.int32, device=dev)
    lid[:mm] = (he_ltok[:mm] | (he_ltok[mm:] << 16)).to(torch.int32)
    tok = torch.zeros((ntiles, tok_stride), dtype=torch.int32, device=dev)
    tok[u_tile, ltok_u] = (uniq - u_tile * n_tokens).to(torch.int32)
    nsub = (counts + row_cap - 1) // row_cap
    n_rows = int(nsub.sum())
    first_row_u = torch.cumsum(nsub, 0) - nsub
    row_u = torch.repeat_interleave(torch.arange(U, **i64), nsub)
    sub = torch.arange(n_rows, **i64) - first_row_u[row_u]
    g_start = torch.cumsum(counts, 0) - counts
    row_tile0 = u_tile[row_u]
    row_len0 = torch.clamp(counts[row_u] - row_cap * sub, max=row_cap)
    # longest rows first inside each tile: the 32 rows a warp sums have (nearly) equal trip counts
    srt = torch.argsort(row_tile0 * 64 + (63 - row_len0), stable=True)
    rank = torch.empty_like(srt); rank[srt] = torch.arange(n_rows, **i64)     # old row id -> sorted position
    row_tile, row_len, row_ltok = row_tile0[srt], row_len0[srt], ltok_u[row_u][srt]
    nrow = torch.bincount(row_tile, minlength=ntiles)
    row_first = torch.cumsum(nrow, 0) - nrow
    r_local = torch.arange(n_rows, **i64) - row_first[row_tile]
    if int(nrow.max()) > rows_stride:
        raise _lib.CfmmError("blocked layout: row table overflow (library/builder mismatch)")
    # g positions: rows are contiguous runs of the tile's row-ordered array, in sorted-row order
    cs = torch.cumsum(row_len, 0) - row_len
    row_start = cs - cs[row_first][row_tile]                                    # per tile
    he_o = torch.arange(2 * mm, **i64) - g_start[inv]                           # offset inside its token group
    he_row = rank[first_row_u[inv] + he_o // row_cap]                           # sorted row id of each half-edge
    he_pos = torch.empty(2 * mm, **i64)
    he_pos[perm] = row_start[he_row] + he_o % row_cap                           # back to (pool, slot) order
    pos = torch.zeros(M, dtype=torch.int32, device=dev)
    # padding pools (last tile) write their zero flows to slots past the real entries of that tile
    pad_base = 2 * (mm - (ntiles - 1) * P)
    if M > mm:
        padl = torch.arange(M - mm, **i64)
        pos[mm:] = ((pad_base + 2 * padl) | ((pad_base + 2 * padl + 1) << 16)).to(torch.int32)
    pos[:mm] = (he_pos[:mm] | (he_pos[mm:] << 16)).to(torch.int32)
    rows = torch.zeros((ntiles, rows_stride), dtype=torch.int32, device=dev)
    word = row_start | (row_len << 16) | (row_ltok << 22)          # start:16 | len:6 | ltok:10 (may set bit 31)
    rows[row_tile, r_local] = torch.where(word >= 2 ** 31, word - 2 ** 32, word).to(torch.int32)
    desc = torch.stack([ntok, nrow, torch.zeros_like(ntok), torch.zeros_like(ntok)], 1).to(torch.int32).contiguous()
    tables = dict(n_tiles=ntiles, M=M, lid=lid, tok=tok, pos=pos, rows=rows, desc=desc,
                  rows_per_pool=n_rows / mm, tok_per_tile=float(ntok.double().mean()))
    return order, residual, tables


class BlockedBucket:
    """Constant-product pools in the token-blocked layout (HBM-bound kind; no per-pool atomics)."""
    kind = _lib.KIND_PRODUCT
    arity = 2
    blocked = True

    def __init__(self, hp: HostPools, spec, device, lib):
        self.spec = spec
        P, rows_stride, tok_stride, row_cap = blocked_layout_info(lib)
        f64 = dict(dtype=torch.float64, device=device)
        self.theta_bar = None
        self.delta = self.lam = self.hcoef = self.hmask = None
        self._device = device
        self._sel = self._off = None
        if spec.identity and int(hp.pool_ptr[-1]) == 2 * hp.m and self._build_native(hp, spec, device, lib, P, rows_stride, tok_stride):
            return                              # the whole layout was built by three launches of csrc/cfmm_layout.cu
        if spec.identity:                       # raw arrays go up as they are; all reordering happens on the GPU
            lo, hi = spec.lo, spec.hi            # a rank's shard uploads only its own slice of the (pinned) host arrays
            R = torch.from_numpy(hp.reserves[2 * lo:2 * hi]).to(device, non_blocking=True).view(-1, 2)
            idx = torch.from_numpy(hp.tok_idx[2 * lo:2 * hi]).to(device, non_blocking=True).view(-1, 2).to(torch.int64)
            gam = torch.from_numpy(hp.gamma[lo:hi]).to(device, non_blocking=True)
            if getattr(hp, "_validate_on_device", False):          # same checks as HostPools.validate(), on the GPU
                chk = torch.stack([R.min(), gam.min(), 1.0 - gam.max(), idx.min().double(),
                                   float(hp.n_tokens - 1) - idx.max().double(),
                                   torch.isfinite(R).all().double() - 0.5]).cpu()
                if bool((chk[:2] <= 0).any()) or bool((chk[2:] < 0).any()):
                    raise ValueError("invalid pool data (reserves > 0, fees in (0, 1], token ids in range, finite)")
        else:
            R = torch.as_tensor(np.ascontiguousarray(hp.reserves[spec.off].T), **f64)
            idx = torch.as_tensor(np.ascontiguousarray(hp.tok_idx[spec.off].T).astype(np.int64), device=device)
            gam = torch.as_tensor(np.ascontiguousarray(hp.gamma[spec.sel]), **f64)
        order, residual, t = build_blocked_pairs(idx.t(), hp.n_tokens, P, rows_stride, tok_stride, row_cap)
        self.order = order                               # blocked position -> bucket-local pool index (device)
        self.residual = residual.cpu().numpy() if residual.numel() else np.zeros(0, np.int64)
        self.m = int(order.numel())
        if self.m == 0:
            self.tables = None
            return
        self.tables = t
        self.stride = t["M"]

        def slab(vals, fill):
            out = torch.full((t["M"],), fill, **f64)
            out[:self.m] = vals
            return out
        self.r0 = slab(R[order, 0], 1.0)
        self.r1 = slab(R[order, 1], 1.0)
        self.gamma_inv = slab(1.0 / gam[order], 1.0)
        self.c_blocked = _lib.BlockedPairs(self.m, t["n_tiles"], P, 0, self.r0.data_ptr(), self.r1.data_ptr(),
                                           self.gamma_inv.data_ptr(), t["lid"].data_ptr(), t["pos"].data_ptr(),
                                           t["rows"].data_ptr(), t["tok"].data_ptr(), t["desc"].data_ptr())

    def _build_native(self, hp, spec, device, lib, P, rows_stride, tok_stride) -> bool:
        """cfmm_blocked_build (csrc/cfmm_layout.cu): upload the pools' own arrays (this rank's slice) and build the blocked
        layout on the device in three launches.  False = not applicable (keys would not fit 32 bits, or some tile would
        touch more tokens than a tile may): the caller takes the general torch builder."""
        lo, hi = spec.lo, spec.hi
        m = hi - lo
        if m <= 0:
            return False
        nb = max(1, int(round((m / P) ** 0.5)))
        if nb * nb * hp.n_tokens >= 2 ** 32 or m >= 2 ** 31:
            return False
        T = -(-m // P)
        M = T * P
        f64 = dict(dtype=torch.float64, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        idx = torch.from_numpy(np.ascontiguousarray(hp.tok_idx[2 * lo:2 * hi], np.int32)).to(device, non_blocking=True)
        R = torch.from_numpy(hp.reserves[2 * lo:2 * hi]).to(device, non_blocking=True)
        gam = torch.from_numpy(hp.gamma[lo:hi]).to(device, non_blocking=True)
        slabs = torch.empty((3, M), **f64)
        words = torch.empty((2, M), **i32)
        rows = torch.empty((T, rows_stride), **i32)
        tok = torch.empty((T, tok_stride), **i32)
        desc = torch.empty((T, 4), **i32)
        order = torch.empty(m, **i32)
        status = torch.empty(4, **i32)
        nbytes = int(lib.cfmm_blocked_build_work_bytes(m))
        if nbytes <= 0:
            return False
        work = torch.empty(nbytes, dtype=torch.uint8, device=device)
        cb = _lib.BlockedPairs(m, T, P, 0, slabs[0].data_ptr(), slabs[1].data_ptr(), slabs[2].data_ptr(), words[0].data_ptr(),
                               words[1].data_ptr(), rows.data_ptr(), tok.data_ptr(), desc.data_ptr())
        st = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        rc = lib.cfmm_blocked_build(m, hp.n_tokens, idx.data_ptr(), R.data_ptr(), gam.data_ptr(), C.byref(cb), order.data_ptr(),
                                    status.data_ptr(), work.data_ptr(), nbytes, st)
        if rc == -3:                              # CFMM_E_SIZE: outside the native builder's key range
            return False
        _lib.check(rc, "cfmm_blocked_build")
        st_h = status.cpu()                       # the one synchronisation of the build
        if int(st_h[1]) != 0:
            raise ValueError("invalid pool data (reserves > 0 and finite, fees in (0, 1], two distinct token ids in range)")
        if int(st_h[0]) != 0:                     # a tile touches more tokens than it may: general builder + plain residual bucket
            return False
        self.order = order
        self.residual = np.zeros(0, np.int64)
        self.m = m
        self.stride = M
        self.r0, self.r1, self.gamma_inv = slabs[0], slabs[1], slabs[2]
        self._keep = (idx, R, gam, work)          # the build is asynchronous: its inputs live as long as the bucket
        self.tables = dict(n_tiles=T, M=M, lid=words[0], pos=words[1], rows=rows, tok=tok, desc=desc,
                           rows_per_pool=int(st_h[2]) / m, tok_per_tile=None)
        self.c_blocked = cb
        return True

    # host-side index maps are only needed for read-back / dense assembly: built on first use
    @property
    def sel(self) -> np.ndarray:
        if self._sel is None:
            self._sel = self.spec.sel[self.order.cpu().numpy().astype(np.int64)]
        return self._sel

    @property
    def off(self) -> np.ndarray:
        if self._off is None:
            self._off = self.spec.hp.pool_ptr[self.sel][None, :] + np.arange(2)[:, None]
        return self._off

    def bytes_resident(self) -> int:
        if self.tables is None:
            return 0
        ts = [self.r0, self.r1, self.gamma_inv] + [self.tables[k] for k in ("lid", "tok", "pos", "rows", "desc")]
        return sum(x.numel() * x.element_size() for x in ts)

    def out_struct(self, trades: bool, hess: bool):
        f64 = dict(dtype=torch.float64, device=self._device)
        if trades and self.delta is None:
            self.delta = torch.zeros((2, self.stride), **f64)
            self.lam = torch.zeros((2, self.stride), **f64)
        if hess and self.hcoef is None:
            self.hcoef = torch.zeros(self.stride, **f64)
        return _lib.EvalOut(self.delta.data_ptr() if trades else None, self.lam.data_ptr() if trades else None,
                            self.hcoef.data_ptr() if hess else None, None)


class PeerContext:
    """Symmetric-memory buffers + sequence counters of the NVLink all-reduce kernel (csrc/cfmm_allreduce.cu), one per
    (process, token count): allocation and rendezvous cost milliseconds, so they are paid once, not per
    PoolStore / per solve.  All ranks must issue the same sequence of reductions (they do: every rank runs the same
    outer loop on bit-identical reduced vectors)."""

    def __init__(self, n_tokens: int, device, group):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        n = int(n_tokens)
        if n + 1 > 64 * 256:
            raise _lib.CfmmError("peer all-reduce supports n_tokens < 16384; use NCCL (Comm) beyond that")
        f64 = dict(dtype=torch.float64, device=device)
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.n_tokens = n
        w = self.world
        # receive areas: [3 slots][world sources][n cells of {value, sequence flag}] (zero = "nothing received yet")
        self.sym_acc = symm.empty((3 * w * (n + 1) * 2,), **f64); self.sym_acc.zero_()
        self.sym_y = symm.empty((3 * w * (n + 2) * 2,), **f64); self.sym_y.zero_()      # n-vectors (+2: the persistent solver appends p'Hp, p'Dp)
        self.hdl_acc = symm.rendezvous(self.sym_acc, group)
        self.hdl_y = symm.rendezvous(self.sym_y, group)
        self.red_acc = torch.zeros(n + 1, **f64)
        self.red_y = torch.zeros(n, **f64)
        self.seq = [0, 0]                   # last sequence number used on channel 0 ([psi | arb]) / 1 (n-vectors)
        torch.cuda.synchronize(device)
        dist.barrier(group)

    def next_seq(self, chan: int) -> int:
        self.seq[chan] += 1
        return self.seq[chan]

    def c_struct(self):
        """cfmm_peer_ctx for the native solvers (they advance the sequence numbers; read them back with absorb())"""
        return _lib.PeerCtx(int(self.hdl_acc.buffer_ptrs_dev), int(self.hdl_y.buffer_ptrs_dev), self.rank, self.world,
                            self.seq[0], self.seq[1])

    def absorb(self, c):
        self.seq = [int(c.seq_acc), int(c.seq_vec)]


_PEER_CONTEXTS: Dict[tuple, "PeerContext"] = {}


def peer_context(n_tokens: int, device, group=None) -> "PeerContext":
    """the process-wide PeerContext for this token count (created on first use; collective)"""
    import torch.distributed as dist
    group = group or dist.group.WORLD
    key = (int(n_tokens), id(group), str(torch.device(device)))
    if key not in _PEER_CONTEXTS:
        _PEER_CONTEXTS[key] = PeerContext(n_tokens, torch.device(device), group)
    return _PEER_CONTEXTS[key]


class PoolStore:
    """All pools of one problem (or one rank's shard of them), resident on one GPU."""

    def __init__(self, hp: HostPools, device="cuda", rank: int = 0, world: int = 1, validate: bool = True,
                 layout: str = "blocked"):
        if layout not in ("blocked", "plain"):
            raise ValueError("layout must be 'blocked' or 'plain'")
        if validate:
            if getattr(hp, "_uniform_product", False) and layout == "blocked":
                hp._validate_on_device = True           # checked after the upload, by reductions on the GPU
            else:
                hp.validate()
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CfmmError("PoolStore needs a CUDA device: there is no CPU path in this package")
        if not torch.cuda.is_available():
            raise _lib.CfmmError("no CUDA device visible: the routing kernels have no CPU fallback")
        self.n_tokens = int(hp.n_tokens)
        self.m_total = hp.m
        self.pool_ptr = hp.pool_ptr
        self._tok_idx_host = hp.tok_idx
        self.rank, self.world = rank, world
        self.buckets = []
        for s in split_buckets(hp, rank, world):
            if s.kind == _lib.KIND_PRODUCT and layout == "blocked" and s.m > 0:
                bb = BlockedBucket(hp, s, self.device, self.lib)
                if bb.m > 0:
                    self.buckets.append(bb)
                if len(bb.residual):
                    self.buckets.append(DeviceBucket(hp, s.subset(bb.residual), self.device))
            else:
                self.buckets.append(DeviceBucket(hp, s, self.device))
        # the blocked bucket (at most one per store) goes first: its launch clears the ping-pong partner buffer
        self.buckets.sort(key=lambda b: 0 if getattr(b, "blocked", False) else 1)
        self._blocked_first = bool(self.buckets) and getattr(self.buckets[0], "blocked", False)
        assert sum(1 for b in self.buckets if getattr(b, "blocked", False)) <= 1
        self.m_local = sum(b.m for b in self.buckets)
        self.has_sum = bool(np.any(hp.kind == KIND_SUM_HOST))
        self.has_geomean = any(b.kind == _lib.KIND_GEOMEAN for b in self.buckets)
        f64 = dict(dtype=torch.float64, device=self.device)
        # [psi | arb] (the one all-reduced buffer) and y are ping-ponged: a blocked launch clears the buffer of the
        # NEXT call, so steady-state evaluations need no memset node
        self._acc2 = torch.zeros((2, self.n_tokens + 1), **f64)
        self._y2 = torch.zeros((2, self.n_tokens), **f64)
        self._acc_i = 0
        self._y_i = 0
        self._move = torch.zeros(1, **f64)
        self.evals = 0
        self.hvps = 0

    # -- multi-GPU: peer-memory all-reduce -----------------------------------------------------------
    def enable_peer_allreduce(self, group=None):
        """Pool-sharded stores (world > 1): finish every evaluate()/hvp()/hess_diag() with the NVLink all-reduce kernel
        cfmm_allreduce_ll (PDL-chained behind the pool kernels: every rank pushes {value, seq} cells into the peers'
        receive areas, one NVLink one-way trip) instead of returning a partial for NCCL.  The buffers live in torch
        symmetric memory and are created ONCE per process and token count (peer_context()); every store of the
        process shares them.  Collective: every rank must call it, in the same order."""
        self._peer = peer_context(self.n_tokens, self.device, group)
        self.reduces_internally = True

    def _peer_reduce(self, chan, local, n, st):
        """all-reduce `local` (n doubles, this rank's partial) over the peer context; chan 0 = [psi | arb], 1 = n-vectors"""
        p = self._peer
        seq = p.next_seq(chan)
        hdl, out = (p.hdl_acc, p.red_acc) if chan == 0 else (p.hdl_y, p.red_y)
        _lib.check(self.lib.cfmm_allreduce_ll(local.data_ptr(), int(hdl.buffer_ptrs_dev), p.rank, p.world, n,
                                              (seq % 3) * p.world * n, n, out.data_ptr(), seq, st), "cfmm_allreduce_ll")
        return out

    # -- helpers -------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def bytes_resident(self) -> int:
        return sum(b.bytes_resident() for b in self.buckets)

    def algorithmic_bytes_per_eval(self) -> int:
        """32 B per 2-token pool, 28k+12 per weighted pool, + nu, psi, arb."""
        n = 0
        for b in self.buckets:
            n += b.m * (28 * b.arity + 12 if b.kind == _lib.KIND_GEOMEAN else 48 if b.kind == _lib.KIND_BOUNDED else 32)
        return n + 16 * self.n_tokens + 8

    # -- the hot path --------------------------------------------------------------------------
    def evaluate(self, nu: torch.Tensor, eps: float = 0.0, trades: bool = False, hess: bool = False, reduce: bool = True):
        """psi(nu) (n_tokens) and arb(nu) (1) for this rank's pools, as views into one (n+1) buffer (all-reduced over
        the peer context when enable_peer_allreduce() was called, unless reduce=False: this rank's partial)."""
        st = self._stream()
        acc = self._acc2[self._acc_i]
        nxt = self._acc2[self._acc_i ^ 1]
        self._acc_i ^= 1
        if not self._blocked_first:       # otherwise the previous blocked launch already cleared `acc`
            _lib.check(self.lib.cfmm_zero(acc.data_ptr(), acc.numel() * 8, st), "cfmm_zero")
        lognu = torch.log(nu) if self.has_geomean else None
        for b in self.buckets:
            out = b.out_struct(trades, hess) if (trades or hess) else None
            if getattr(b, "blocked", False):
                rc = self.lib.cfmm_blocked_eval(C.byref(b.c_blocked), self.n_tokens, nu.data_ptr(), acc.data_ptr(),
                                                acc.data_ptr() + 8 * self.n_tokens,
                                                C.byref(out) if out is not None else None,
                                                nxt.data_ptr(), nxt.numel(), st)
                _lib.check(rc, "cfmm_blocked_eval")
                continue
            rc = self.lib.cfmm_arb_eval(C.byref(b.c_bucket), self.n_tokens, nu.data_ptr(),
                                        lognu.data_ptr() if lognu is not None else None, float(eps),
                                        acc.data_ptr(), acc.data_ptr() + 8 * self.n_tokens,
                                        C.byref(out) if out is not None else None, st)
            _lib.check(rc, "cfmm_arb_eval")
        self.evals += 1
        if reduce and getattr(self, "reduces_internally", False):
            return self._peer_reduce(0, acc, self.n_tokens + 1, st)
        return acc

    def hvp(self, vt: torch.Tensor) -> torch.Tensor:
        st = self._stream()
        y = self._y2[self._y_i]
        ynxt = self._y2[self._y_i ^ 1]
        self._y_i ^= 1
        if not self._blocked_first:
            _lib.check(self.lib.cfmm_zero(y.data_ptr(), y.numel() * 8, st), "cfmm_zero")
        for b in self.buckets:
            if getattr(b, "blocked", False):
                _lib.check(self.lib.cfmm_blocked_hvp(C.byref(b.c_blocked), self.n_tokens, b.hcoef.data_ptr(),
                                                     vt.data_ptr(), y.data_ptr(), ynxt.data_ptr(), st),
                           "cfmm_blocked_hvp")
                continue
            rc = self.lib.cfmm_hvp(C.byref(b.c_bucket), self.n_tokens, b.hcoef.data_ptr(),
                                   b.hmask.data_ptr(), vt.data_ptr(), y.data_ptr(), st)
            _lib.check(rc, "cfmm_hvp")
        self.hvps += 1
        if getattr(self, "reduces_internally", False):
            return self._peer_reduce(1, y, self.n_tokens, st)
        return y

    def hess_diag(self) -> torch.Tensor:
        st = self._stream()
        d = torch.zeros(self.n_tokens, dtype=torch.float64, device=self.device)
        for b in self.buckets:
            if getattr(b, "blocked", False):
                _lib.check(self.lib.cfmm_blocked_diag(C.byref(b.c_blocked), self.n_tokens, b.hcoef.data_ptr(),
                                                      d.data_ptr(), st), "cfmm_blocked_diag")
                continue
            _lib.check(self.lib.cfmm_hess_diag(C.byref(b.c_bucket), self.n_tokens, b.hcoef.data_ptr(),
                                               b.hmask.data_ptr(), d.data_ptr(), st), "cfmm_hess_diag")
        if getattr(self, "reduces_internally", False):
            return self._peer_reduce(1, d, self.n_tokens, st).clone()
        return d

    def hess_dense(self) -> torch.Tensor:
        st = self._stream()
        H = torch.zeros((self.n_tokens, self.n_tokens), dtype=torch.float64, device=self.device)
        for b in self.buckets:
            if getattr(b, "blocked", False):
                _lib.check(self.lib.cfmm_blocked_dense(C.byref(b.c_blocked), self.n_tokens, b.hcoef.data_ptr(), H.data_ptr(), st),
                           "cfmm_blocked_dense")
                continue
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
