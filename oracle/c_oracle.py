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
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def num_threads():
    return int(load().oracle_num_threads())


def eval_pairs(idx, R, gamma, n_tokens, nu, want_trades=False):
    """idx (m,2) int32, R (m,2) f64, gamma (m,) -> psi (n,), arb, [delta (m,2), lam (m,2)]"""
    lib = load()
    idx = np.ascontiguousarray(idx, np.int32); R = np.ascontiguousarray(R, np.float64)
    gamma = np.ascontiguousarray(gamma, np.float64); nu = np.ascontiguousarray(nu, np.float64)
    m = len(gamma)
    psi = np.empty(n_tokens); arb = np.zeros(1)
    d = np.empty((m, 2)) if want_trades else None
    l = np.empty((m, 2)) if want_trades else None
    rc = lib.oracle_eval_pairs(m, idx.ctypes.data, R.ctypes.data, gamma.ctypes.data, n_tokens, nu.ctypes.data,
                               psi.ctypes.data, arb.ctypes.data, d.ctypes.data if want_trades else None,
                               l.ctypes.data if want_trades else None)
    if rc:
        raise MemoryError("oracle_eval_pairs")
    return (psi, float(arb[0]), d, l) if want_trades else (psi, float(arb[0]))
