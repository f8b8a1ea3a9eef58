"""BASELINE.json configs[0..3] end to end on the GPU (configs[4] is bench.py): value, certificate, wall time."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfmm_routing_code_b200 as cf
from cfmm_routing_code_b200 import instances as I

def run(name, hp, util, **kw):
    for rep in range(2):                               # second run: warm library / allocator
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = cf.solve_pools(hp, util, tol=1e-6, want_trades=False, **kw)
        torch.cuda.synchronize(); wall = time.perf_counter() - t0
    print(f"{name:58s} value {r.value:.9g} gap {r.gap:+.1e} infeas {r.primal_infeas:.1e} {r.status} "
          f"iters {r.iters} evals {r.evals} hvps {r.hvps} wall {1e3*wall:.1f} ms (solver {1e3*r.wall_s:.1f} ms)", flush=True)

d = I.arbitrage_instance()
run("cfg1 arbitrage.py (5 pools, 4 tokens)", cf.HostPools.from_lists(4, d["local_indices"], d["reserves"], d["fees"], d["kinds"], d["weights"]), cf.Arbitrage(d["market_value"]))
s = I.synth_const_product(10_000, 256, seed=0)
run("cfg2 10k constant-product pools, 256 tokens", cf.HostPools.from_pairs(256, s["idx"], s["reserves"], s["gamma"]), cf.Arbitrage(s["prices"]))
s = I.synth_mixed(100_000, 1000, seed=1)
hp = cf.HostPools(1000, s["pool_ptr"], s["tok_idx"], s["reserves"], s["weights"], s["gamma"], s["kind"])
run("cfg3 100k mixed pools (60% product, 30% weighted, 10% sum), 1k tok", hp, cf.Arbitrage(s["prices"]))
s = I.synth_mixed(100_000, 1000, seed=2)
hp = cf.HostPools(1000, s["pool_ptr"], s["tok_idx"], s["reserves"], s["weights"], s["gamma"], s["kind"])
basket = I.synth_basket(1000, s["prices"], seed=2)
run("cfg4 liquidation of a 16-token basket over 100k mixed pools", hp, cf.Liquidate(0, basket), nu0=s["prices"] / s["prices"][0])
s = I.synth_const_product(1_000_000, 4096, seed=3)
run("cfg5 1M constant-product pools, 4096 tokens (native solver)", cf.HostPools.from_pairs(4096, s["idx"], s["reserves"], s["gamma"]).pin_memory(), cf.Arbitrage(s["prices"]))
