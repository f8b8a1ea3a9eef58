"""ncu target: a handful of eval launches in the scatter modes given on the command line."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfmm_routing_code_b200 as cf
from cfmm_routing_code_b200 import instances as I
m, n = 1_000_000, 4096
stores = []
for k in range(6):
    s = I.synth_const_product(m, n, seed=3 + k)
    stores.append((cf.PoolStore(cf.HostPools.from_pairs(n, s["idx"], s["reserves"], s["gamma"]), validate=False), s))
nu = torch.as_tensor(stores[0][1]["prices"] * np.exp(0.01 * np.random.default_rng(0).standard_normal(n)),
                     dtype=torch.float64, device="cuda")
lib = stores[0][0].lib
for mode in [0]:
    for k in range(6):
        stores[k][0].evaluate(nu)
    torch.cuda.synchronize()
