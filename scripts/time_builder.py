import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfmm_routing_code_b200 as cf
from cfmm_routing_code_b200 import instances as I, pools as PL
m, n = 1_000_000, 4096
s = I.synth_const_product(m, n, seed=3)
hp = cf.HostPools.from_pairs(n, s["idx"], s["reserves"], s["gamma"])
for layout in ("plain", "blocked", "blocked", "plain"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = cf.PoolStore(hp, validate=False, layout=layout)
    torch.cuda.synchronize(); print(layout, "PoolStore build %.1f ms" % (1e3 * (time.perf_counter() - t0)))
lib = st.lib
P, rs, ts, rc, es = PL.blocked_layout_info(lib)
idx = torch.as_tensor(hp.tok_idx.reshape(-1, 2).T.astype(np.int64).copy(), device="cuda")
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    PL.build_blocked_pairs(idx, n, P, rs, ts, rc, es)
    torch.cuda.synchronize(); print("build_blocked_pairs alone %.1f ms" % (1e3 * (time.perf_counter() - t0)))
