"""Small workload that touches every kernel family once, for compute-sanitizer (memcheck / racecheck / synccheck):
blocked evaluation (TMA ring + mbarriers + programmatic dependent launch), register-fed hvp / diag, the plain TMA pair
kernel, weighted pools, constant-sum multipliers, the persistent solver (grid barriers, slice phases) and the per-thread
batch solver.  Sizes are small: the sanitizer slows kernels down 10-100x."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cfmm_routing_code_b200 as cf
from cfmm_routing_code_b200 import instances as I
import helpers as H

F64 = dict(dtype=torch.float64, device="cuda")
hp, s = H.cp_host_pools(6000, 300, seed=1)                 # 7 tiles: several CTAs, ragged last tile
st = cf.PoolStore(hp)
nu = torch.as_tensor(H.random_prices(s["prices"], 1), **F64)
for k in range(4):                                          # back-to-back: PDL chaining + ping-pong clears
    acc = st.evaluate(nu * (1 + 1e-3 * k), trades=(k == 3), hess=True).clone()
v = torch.randn(300, **F64)
y = st.hvp(v).clone(); d = st.hess_diag().clone()
stp = cf.PoolStore(hp, layout="plain")
accp = stp.evaluate(nu * (1 + 3e-3), hess=True).clone()
assert float((acc - accp).abs().max()) < 1e-6 * float(accp.abs().max())
hm, sm = H.mixed_host_pools(3000, 60, seed=2)
stm = cf.PoolStore(hm)
stm.evaluate(torch.as_tensor(H.random_prices(sm["prices"], 3, 0.03), **F64), 1e-3, trades=True, hess=True)
stm.update_multipliers(); stm.hess_dense(); stm.hvp(torch.randn(60, **F64)); stm.hess_diag()
for impl in ("persist", "hostloop"):
    r = cf.solve_pools(hp, cf.Arbitrage(s["prices"]), tol=1e-8, store=st, native=impl, want_trades=False)
    assert r.status == "optimal", (impl, r.status)
    print(impl, r.status, r.iters, r.evals, r.hvps, r.value)
dd = I.two_asset_instance()
rs = cf.solve_sweep(dd["local_indices"], dd["reserves"], dd["fees"], dd["kinds"], dd["weights"],
                    [cf.Swap(0, 2, t) for t in dd["amounts"][::5]], tol=1e-9)
assert all(r.status == "optimal" for r in rs)
torch.cuda.synchronize()
print("sanitize workload ok")
