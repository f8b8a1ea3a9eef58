"""Test-only: builds tests/host_harness/small_host.cpp (the per-problem solver of csrc/cfmm_small.cuh compiled for the
host) and calls it, so the control flow the CUDA kernel runs per thread can be checked without a GPU.  The product never
loads this library."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_harness", "small_host.cpp")
HDR = os.path.join(HERE, "..", "cfmm_routing_code_b200", "csrc", "cfmm_small.cuh")
LIB = os.path.join(HERE, "_build", "libsmall_host.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        stale = (not os.path.exists(LIB)) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR))
        if stale:
            os.makedirs(os.path.dirname(LIB), exist_ok=True)
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Werror", "-o", LIB, SRC],
                           check=True)
        _lib = C.CDLL(LIB)
        _lib.small_host_solve.argtypes = ([C.c_int, C.c_longlong] + [C.c_void_p] * 7 + [C.c_int] + [C.c_void_p] * 9
                                          + [C.c_longlong, C.c_double, C.c_int])
    return _lib


def solve(hp, specs, tol=1e-9, interleave=1, nu0=None, pool_range=None):
    """hp: HostPools; specs: objects with c, a, eq, pinned.  Returns dict(nu, psi, stats, delta, lam)."""
    n, B = hp.n_tokens, len(specs)
    c = np.stack([np.asarray(u.c, float) for u in specs])
    a = np.stack([np.asarray(u.a, float) for u in specs])
    fl = np.stack([np.asarray(u.eq, np.uint8) | (np.asarray(u.pinned, np.uint8) << 1) for u in specs])
    nu = np.empty((B, n))
    for p, u in enumerate(specs):
        pos = u.c[u.c > 0]
        nu[p] = np.where(u.c > 0, u.c, np.median(pos) if len(pos) else 1.0) if nu0 is None else nu0[p]
    return solve_raw(hp, c, a, fl, nu, pool_range, tol, interleave)


def solve_raw(hp, c, a, fl, nu, pool_range=None, tol=1e-9, interleave=1):
    """the arrays cfmm_batch_solve takes (see include/cfmm_b200.h: cfmm_csr_pools, cfmm_batch), on the host"""
    lib = load()
    n, B, nnz = hp.n_tokens, len(c), len(hp.tok_idx)
    c = np.ascontiguousarray(c, np.float64); a = np.ascontiguousarray(a, np.float64)
    fl = np.ascontiguousarray(fl, np.uint8); nu = np.ascontiguousarray(nu, np.float64).copy()
    psi = np.zeros((B, n)); st = np.zeros((B, 8))
    shared = pool_range is None
    d = np.zeros((B if shared else 1, nnz)); l = np.zeros_like(d)
    slot_kind = np.repeat(np.asarray(hp.kind), np.diff(hp.pool_ptr))
    logrw = np.log(np.maximum(hp.reserves, 1e-300) / np.where(slot_kind == 0, hp.weights, 1.0))
    keep = [np.ascontiguousarray(hp.pool_ptr, np.int64), np.ascontiguousarray(hp.tok_idx, np.int32),
            np.ascontiguousarray(hp.reserves, np.float64), np.ascontiguousarray(hp.weights, np.float64),
            np.ascontiguousarray(logrw), np.ascontiguousarray(hp.gamma, np.float64),
            np.ascontiguousarray(hp.kind, np.uint8)]
    pr = None if shared else np.ascontiguousarray(pool_range, np.int64)
    p_ = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)
    lib.small_host_solve(n, hp.m, *[p_(k) for k in keep], B, p_(pr), p_(c), p_(a), p_(fl), p_(nu), p_(psi), p_(st),
                         p_(d), p_(l), nnz if shared else 0, tol, interleave)
    return dict(nu=nu, psi=psi, stats=st, delta=d, lam=l)
