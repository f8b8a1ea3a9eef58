"""ctypes wrapper of oracle/cfmm_oracle_c.c (pthreads dual evaluation of constant-product pools).  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def load():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "libcfmm_oracle_c.so")
        src = os.path.join(HERE, "cfmm_oracle_c.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.run(["make", "-C", HERE, "-s"], check=True)
        _lib = C.CDLL(so)
        _lib.oracle_eval_pairs.restype = C.c_int
        _lib.oracle_eval_pairs.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_hess_pairs.restype = C.c_int
        _lib.oracle_hess_pairs.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int]
        _lib.oracle_num_threads.restype = C.c_int
        _lib.oracle_solve_pairs.restype = C.c_int
        _lib.oracle_solve_pairs.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_int32,
                                            C.POINTER(SolveResult)]
    return _lib


class SolveResult(C.Structure):
    _fields_ = [("dual_value", C.c_double), ("primal_value", C.c_double), ("gap", C.c_double),
                ("primal_infeas", C.c_double), ("err", C.c_double), ("wall_s", C.c_double), ("iters", C.c_int32),
                ("evals", C.c_int32), ("hvps", C.c_int32), ("status", C.c_int32)]


def solve_pairs(idx, R, gamma, n_tokens, c, a=None, eq=None, pinned=None, nu0=None, tol=1e-8, max_iter=100, cg_max=200):
    """oracle_solve_pairs: the whole dual solve of a constant-product problem in C (projected Newton-PCG, persistent
    pthread pool): the CPU arm of bench.py.  Returns (nu, psi, SolveResult)."""
    lib = load()
    idx = np.ascontiguousarray(idx, np.int32); R = np.ascontiguousarray(R, np.float64)
    gamma = np.ascontiguousarray(gamma, np.float64)
    n = int(n_tokens)
    c = np.ascontiguousarray(c, np.float64)
    a = np.zeros(n) if a is None else np.ascontiguousarray(a, np.float64)
    eq = np.zeros(n, np.uint8) if eq is None else np.ascontiguousarray(eq, np.uint8)
    pinned = np.zeros(n, np.uint8) if pinned is None else np.ascontiguousarray(pinned, np.uint8)
    if nu0 is None:
        pos = c[c > 0]
        nu = np.where(c > 0, c, np.median(pos) if len(pos) else 1.0).astype(np.float64)
    else:
        nu = np.array(nu0, np.float64)
    psi = np.empty(n)
    res = SolveResult()
    rc = lib.oracle_solve_pairs(len(gamma), idx.ctypes.data, R.ctypes.data, gamma.ctypes.data, n, c.ctypes.data,
                                a.ctypes.data, eq.ctypes.data, pinned.ctypes.data, nu.ctypes.data, psi.ctypes.data,
                                float(tol), int(max_iter), int(cg_max), C.byref(res))
    if rc:
        raise MemoryError("oracle_solve_pairs")
    return nu, psi, res


def num_threads():
    return int(load().oracle_num_threads())


def set_threads(n):
    load().oracle_set_threads(int(n))


def autotune_threads(idx, R, gamma, n_tokens, nu, candidates=(4, 8, 16, 32, 64, 128, 256)):
    """pick the thread count with the best evaluation throughput on this host (more is not better: every thread
    owns a private psi array and the scatter is memory bound)"""
    import os
    import time
    ncpu = os.cpu_count() or 1
    best, best_t = None, float("inf")
    for c in sorted({min(c, ncpu) for c in candidates}):
        set_threads(c)
        eval_pairs(idx, R, gamma, n_tokens, nu)
        t0 = time.perf_counter()
        for _ in range(3):
            eval_pairs(idx, R, gamma, n_tokens, nu)
        dt = (time.perf_counter() - t0) / 3
        if dt < best_t:
            best, best_t = c, dt
    set_threads(best)
    return best


def eval_pairs(idx, R, gamma, n_tokens, nu, want_trades=False, hcoef=None):
    """idx (m,2) int32, R (m,2) f64, gamma (m,) -> psi (n,), arb, [delta (m,2), lam (m,2)]; hcoef (m,) filled if given"""
    lib = load()
    idx = np.ascontiguousarray(idx, np.int32); R = np.ascontiguousarray(R, np.float64)
    gamma = np.ascontiguousarray(gamma, np.float64); nu = np.ascontiguousarray(nu, np.float64)
    m = len(gamma)
    psi = np.empty(n_tokens); arb = np.zeros(1)
    d = np.empty((m, 2)) if want_trades else None
    l = np.empty((m, 2)) if want_trades else None
    rc = lib.oracle_eval_pairs(m, idx.ctypes.data, R.ctypes.data, gamma.ctypes.data, n_tokens, nu.ctypes.data,
                               psi.ctypes.data, arb.ctypes.data, d.ctypes.data if want_trades else None,
                               l.ctypes.data if want_trades else None, hcoef.ctypes.data if hcoef is not None else None)
    if rc:
        raise MemoryError("oracle_eval_pairs")
    return (psi, float(arb[0]), d, l) if want_trades else (psi, float(arb[0]))


def hess_pairs(idx, hcoef, n_tokens, vt=None):
    """y = Hs vt (vt given) or diag(Hs) (vt None) for constant-product pools, log-price coordinates"""
    lib = load()
    out = np.empty(n_tokens)
    vtc = np.ascontiguousarray(vt, np.float64) if vt is not None else None
    rc = lib.oracle_hess_pairs(len(hcoef), idx.ctypes.data, hcoef.ctypes.data, n_tokens,
                               vtc.ctypes.data if vtc is not None else None, out.ctypes.data, 1 if vt is not None else 2)
    if rc:
        raise MemoryError("oracle_hess_pairs")
    return out


class CpuPairsEvaluator:
    """The solver's evaluator protocol on the host cores (constant-product pools only): what bench.py's
    `--impl reference` leg solves with.  CPU torch tensors in/out, C + pthreads underneath."""

    def __init__(self, n_tokens, idx, R, gamma):
        import torch
        self.torch = torch
        self.n_tokens = int(n_tokens)
        self.idx = np.ascontiguousarray(idx, np.int32); self.R = np.ascontiguousarray(R, np.float64)
        self.gamma = np.ascontiguousarray(gamma, np.float64)
        self.hcoef = np.zeros(len(self.gamma))
        self.has_sum = False
        self.device = torch.device("cpu")
        self.evals = 0
        self.hvps = 0

    def evaluate(self, nu, eps=0.0, trades=False, hess=False):
        psi, arb = eval_pairs(self.idx, self.R, self.gamma, self.n_tokens, nu.numpy(), hcoef=self.hcoef if hess else None)
        self.evals += 1
        return self.torch.as_tensor(np.concatenate([psi, [arb]]))

    def hvp(self, vt):
        self.hvps += 1
        return self.torch.as_tensor(hess_pairs(self.idx, self.hcoef, self.n_tokens, vt.numpy()))

    def hess_diag(self):
        return self.torch.as_tensor(hess_pairs(self.idx, self.hcoef, self.n_tokens))

    def reset_multipliers(self):
        pass
