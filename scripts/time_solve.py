"""Times the 1M-pool solve (BASELINE configs[4]) through both native loops, and the stages of the e2e call."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfmm_routing_code_b200 as cf
from cfmm_routing_code_b200 import instances as I

m, n = int(os.environ.get("M_POOLS", 1_000_000)), 4096
s = I.synth_const_product(m, n, seed=3)
hp = cf.HostPools.from_pairs(n, s["idx"], s["reserves"], s["gamma"]).pin_memory()
util = cf.Arbitrage(s["prices"])
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); st = cf.PoolStore(hp, validate=False); torch.cuda.synchronize(); tb = time.perf_counter() - t0
print(f"PoolStore build (upload + layout) {1e3*tb:.2f} ms")
import ctypes
for impl, mode in (("persist", 0), ("hostloop", 0)):
    for tol in (1e-6, 1e-9):
        ws = []
        for rep in range(5):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = cf.solve_pools(hp, util, tol=tol, store=st, native=impl, want_trades=False)
            torch.cuda.synchronize(); ws.append(time.perf_counter() - t0)
        name = impl if impl == "hostloop" else ("persist/dist" if mode == 0 else "persist/boss")
        print(f"{name:12s} tol {tol:.0e}: {r.status} iters {r.iters} evals {r.evals} hvps {r.hvps} gap {r.gap:+.1e} infeas {r.primal_infeas:.1e} "
              f"solve wall min {1e3*min(ws):.3f} ms median {1e3*sorted(ws)[2]:.3f} ms (solver-only {1e3*r.wall_s:.3f} ms) value {r.value:.9g}", flush=True)
        if impl == "persist":
            prof = (ctypes.c_int64 * 16)()
            st.lib.cfmm_persist_last_profile(prof)
            us = [x / 1965.0 for x in prof]            # SM cycles -> us at 1965 MHz
            if mode == 0:
                print(f"   CTA0 profile (us): pass eval {us[0]:.0f} hvp {us[1]:.0f} diag {us[2]:.0f} | barrier A {us[3]:.0f} | slice phase {us[4]:.0f} | "
                      f"barrier B {us[5]:.0f} | decide {us[6]:.0f}")
            else:
                print(f"   CTA0 profile (us): pass eval {us[0]:.0f} hvp {us[1]:.0f} diag {us[2]:.0f} | wait-grid {us[3]:.0f} | "
                      f"algebra after eval {us[4]:.0f} hvp {us[5]:.0f} diag {us[6]:.0f}")
ws = []
for rep in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = cf.solve_pools(hp, util, tol=1e-6, want_trades=False)
    torch.cuda.synchronize(); ws.append(time.perf_counter() - t0)
print(f"e2e solve_pools from pinned host arrays: min {1e3*min(ws):.3f} ms median {1e3*sorted(ws)[2]:.3f} ms")
