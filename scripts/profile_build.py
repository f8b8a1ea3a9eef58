"""Where does PoolStore's build time go (upload + layout build of the 1M-pool instance)?  torch profiler table."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfmm_routing_code_b200 as cf
from cfmm_routing_code_b200 import instances as I
from torch.profiler import profile, ProfilerActivity

s = I.synth_const_product(1_000_000, 4096, seed=3)
hp = cf.HostPools.from_pairs(4096, s["idx"], s["reserves"], s["gamma"]).pin_memory()
for _ in range(3):
    st = cf.PoolStore(hp, validate=False); torch.cuda.synchronize()
ts = []
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter(); st = cf.PoolStore(hp, validate=False); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("build wall ms", [round(1e3 * t, 2) for t in ts])
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    st = cf.PoolStore(hp, validate=False); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
print("cuda kernels:", len(ev), "total cuda us", sum(e.cuda_time for e in ev) if ev else None)
