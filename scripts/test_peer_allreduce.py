"""2-GPU check (torchrun): pool-sharded evaluation with the fused NVLink one-shot all-reduce vs NCCL, then timing."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cfmm_routing_code_b200 as cf
from cfmm_routing_code_b200 import instances as I

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
m, n = 1_000_000, 4096
s = I.synth_const_product(m, n, seed=3)
hp = cf.HostPools.from_pairs(n, s["idx"], s["reserves"], s["gamma"])
nu = torch.as_tensor(s["prices"] * np.exp(0.01 * np.random.default_rng(0).standard_normal(n)), dtype=torch.float64, device=dev)
st_nccl = cf.PoolStore(hp, device=dev, rank=rank, world=world, validate=False)
st_peer = cf.PoolStore(hp, device=dev, rank=rank, world=world, validate=False)
st_peer.enable_peer_allreduce(protocol=os.environ.get("PEER_PROTOCOL", "ll"))
print(rank, "fused:", getattr(st_peer, "_peer_fused", None), flush=True) if rank == 0 else None
ok = True
for it in range(7):                      # > 3 rounds: exercises the slot rotation
    nui = nu * (1 + 0.001 * it)
    a = st_nccl.evaluate(nui).clone(); dist.all_reduce(a)
    b = st_peer.evaluate(nui).clone()
    err = float((a - b).abs().max() / a.abs().max())
    gathered = [torch.zeros_like(b) for _ in range(world)]; dist.all_gather(gathered, b)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    ok &= err < 1e-12 and same
    if rank == 0: print(f"round {it}: rel diff vs NCCL {err:.2e}  bit-identical across ranks {same}", flush=True)
st_nccl.evaluate(nu, hess=True); st_peer.evaluate(nu, hess=True)
v = torch.randn(n, dtype=torch.float64, device=dev); dist.broadcast(v, 0)
for it in range(4):
    a = st_nccl.hvp(v).clone(); dist.all_reduce(a)
    b = st_peer.hvp(v).clone()
    err = float((a - b).abs().max() / a.abs().max()); ok &= err < 1e-12
    if rank == 0: print(f"hvp round {it}: rel diff {err:.2e}", flush=True)
# full solve through both paths
r1 = cf.solve_pools(hp, cf.Arbitrage(s["prices"]), tol=1e-6, store=st_nccl, want_trades=False)
r2 = cf.solve_pools(hp, cf.Arbitrage(s["prices"]), tol=1e-6, store=st_peer, want_trades=False)
if rank == 0: print("solve nccl", r1.status, r1.value, r1.wall_s, "| peer", r2.status, r2.value, r2.wall_s, flush=True)
ok &= abs(r1.value - r2.value) <= 1e-9 * abs(r1.value)

def timeit(fn, reps=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
t_n = timeit(lambda: dist.all_reduce(st_nccl.evaluate(nu)))
t_p = timeit(lambda: st_peer.evaluate(nu))
if rank == 0: print(os.environ.get("PEER_PROTOCOL", "ll"), f"eager step: nccl {t_n:.1f} us   peer one-shot {t_p:.1f} us   ALL OK={ok}", flush=True)
dist.barrier()
os._exit(0 if ok else 1)
