"""Evaluation of BASELINE configs[2] (100k mixed pools, 1k tokens) a few times: the workload ncu profiles for the
non-product kernels (k_eval_geomean<K>, k_eval_pair_tma<SUM>, k_blocked on the product share)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfmm_routing_code_b200 as cf
from cfmm_routing_code_b200 import instances as I

s = I.synth_mixed(100_000, 1000, seed=1)
hp = cf.HostPools(1000, s["pool_ptr"], s["tok_idx"], s["reserves"], s["weights"], s["gamma"], s["kind"])
st = cf.PoolStore(hp)
nu = torch.as_tensor(s["prices"] * np.exp(0.01 * np.random.default_rng(0).standard_normal(1000)), dtype=torch.float64, device="cuda")
for k in range(6):
    st.evaluate(nu * (1 + 1e-3 * k), 1e-3, hess=(k % 2 == 0))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(50):
    st.evaluate(nu, 1e-3, hess=True)
e1.record(); torch.cuda.synchronize()
print("buckets:", [(b.kind, b.arity, b.m) for b in st.buckets])
print(f"mixed evaluation (all buckets, eager launches): {e0.elapsed_time(e1) / 50 * 1e3:.1f} us; algorithmic bytes {st.algorithmic_bytes_per_eval()}")
