"""Ad-hoc GPU microbenchmarks (development tool, not the bench contract)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfmm_routing_code_b200 as cf
from cfmm_routing_code_b200 import instances as I

dev = "cuda"
m, n = 1_000_000, 4096
ninst = 8
stores = []
for k in range(ninst):
    s = I.synth_const_product(m, n, seed=3 + k)
    hp = cf.HostPools.from_pairs(n, s["idx"], s["reserves"], s["gamma"])
    stores.append((cf.PoolStore(hp, validate=False), s))
nus = [torch.as_tensor(s["prices"] * np.exp(0.01 * np.random.default_rng(k).standard_normal(n)),
                       dtype=torch.float64, device=dev) for k, (_, s) in enumerate(stores)]
lib = stores[0][0].lib


def timeit(fn, iters=200, warm=16):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


for mode in (1, 3, 4):
    lib.cfmm_set_scatter_mode(mode)
    for label, kw in (("eval", {}), ("eval+hess", dict(hess=True)), ("eval+trades", dict(trades=True))):
        t_rot = timeit(lambda i: stores[i % ninst][0].evaluate(nus[i % ninst], **kw))
        t_hot = timeit(lambda i: stores[0][0].evaluate(nus[0], **kw))
        print(f"mode {mode} {label:12s} rotating {t_rot*1e6:8.1f} us  {m/t_rot/1e9:7.2f} Gpool/s  "
              f"{32*m/t_rot/1e9:7.0f} GB/s | L2-hot {t_hot*1e6:8.1f} us {32*m/t_hot/1e9:7.0f} GB/s", flush=True)
    st = stores[0][0]
    st.evaluate(nus[0], hess=True)
    v = torch.randn(n, dtype=torch.float64, device=dev)
    t = timeit(lambda i: st.hvp(v))
    print(f"mode {mode} hvp L2-hot {t*1e6:8.1f} us", flush=True)
lib.cfmm_set_scatter_mode(0)
for ls in ("cg",):
    st, s = stores[0]
    t0 = time.perf_counter()
    r = cf.solve_pools(None if False else cf.HostPools.from_pairs(n, s["idx"], s["reserves"], s["gamma"]),
                       cf.Arbitrage(s["prices"]), tol=1e-6, store=st, want_trades=False, linear_solver=ls)
    torch.cuda.synchronize()
    print(f"solve[{ls}] {time.perf_counter()-t0:.4f}s status={r.status} iters={r.iters} evals={r.evals} "
          f"hvps={r.hvps} gap={r.gap:.2e} infeas={r.primal_infeas:.2e} value={r.value:.6f}")
    for h in r.info.history:
        print("   t=%.4f evals=%d err=%.3e" % h)
