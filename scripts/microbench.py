"""Ad-hoc GPU microbenchmarks (development tool, not the bench contract).  CUDA-graph replay removes host overhead."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfmm_routing_code_b200 as cf
from cfmm_routing_code_b200 import instances as I

dev = "cuda"
m, n = int(os.environ.get('M_POOLS', 1_000_000)), 4096
ninst = 8
stores = []
LAYOUT = os.environ.get("LAYOUT", "blocked")
from cfmm_routing_code_b200 import _lib as _L
_L.load().cfmm_set_blocked_config(200 + int(os.environ.get("PDL", "1")))
_L.load().cfmm_set_blocked_config(300 + int(os.environ.get("ROWCAP", "32")))
for k in range(ninst):
    s = I.synth_const_product(m, n, seed=3 + k)
    hp = cf.HostPools.from_pairs(n, s["idx"], s["reserves"], s["gamma"])
    t0 = time.perf_counter()
    stores.append((cf.PoolStore(hp, validate=False, layout=LAYOUT), s))
    torch.cuda.synchronize()
    if k < 2:
        b = stores[-1][0].buckets[0]
        print(f"store build {time.perf_counter()-t0:.3f}s layout={LAYOUT}",
              {k2: v for k2, v in (b.tables or {}).items() if k2 in ("n_tiles", "rows_per_pool", "tok_per_tile")}
              if getattr(b, "blocked", False) else "", flush=True)
nus = [torch.as_tensor(s["prices"] * np.exp(0.01 * np.random.default_rng(k).standard_normal(n)),
                       dtype=torch.float64, device=dev) for k, (_, s) in enumerate(stores)]
lib = stores[0][0].lib


def graph_time(fn, reps=64, replays=5):
    """capture `reps` calls of fn(i) into one CUDA graph, replay, return seconds per call"""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(ninst):
            fn(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * replays) * 1e-3


for mode in [0]:
    for label, kw in (("eval", {}), ("eval+hess", dict(hess=True))):
        t_rot = graph_time(lambda i: stores[i % ninst][0].evaluate(nus[i % ninst], **kw))
        t_hot = graph_time(lambda i: stores[0][0].evaluate(nus[0], **kw))
        print(f"mode {mode:2d} {label:12s} rotating {t_rot*1e6:8.1f} us  {m/t_rot/1e9:7.2f} Gpool/s  "
              f"{32*m/t_rot/1e9:7.0f} GB/s | L2-hot {t_hot*1e6:8.1f} us {32*m/t_hot/1e9:7.0f} GB/s", flush=True)
st = stores[0][0]
st.evaluate(nus[0], hess=True)
v = torch.randn(n, dtype=torch.float64, device=dev)
print(f"hvp L2-hot {graph_time(lambda i: st.hvp(v))*1e6:8.1f} us", flush=True)
