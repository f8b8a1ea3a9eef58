"""The pool-sharded protocols on ONE GPU: `world` "ranks" are `world` host threads, each with its own CUDA stream, its own
shard of the pools and its own receive areas -- all on cuda:0.  The kernels cannot tell: they push 16-byte {value, seq}
cells through the peers' receive-area pointers and poll their own, exactly as over NVLink peer memory.  This covers what
the 1-GPU test tier otherwise cannot: the in-kernel exchange of the persistent solver (cfmm_persist.cu) and the LL
all-reduce kernel (cfmm_allreduce.cu) at world = 2, 4 and 8 -- slot rotation, rank-order sums (bit-identical results on
all ranks), the scalar extras, uneven shards.  (The real multi-GPU run is tests/test_multigpu.py.)

The ranks' kernels must be co-resident (they wait for each other), so the problems are small: world x tiles CTAs fit the
GPU at once, and the persistent solver is launched WITHOUT the cooperative attribute here (cfmm_set_persist_cooperative(0):
that guarantee is per kernel, the driver may serialise cooperative grids of different streams; 4-17 CTA grids on an idle
148-SM device are resident anyway).  A rank that cannot start ends by the kernels' ~3 s spin limit with an error, not a
hang."""
import ctypes as C
import threading

import numpy as np
import pytest
import torch

import cfmm_routing_code_b200 as cf
from cfmm_routing_code_b200 import _lib
from cfmm_routing_code_b200.solver import default_nu0
import helpers as H

pytestmark = pytest.mark.gpu

F64 = dict(dtype=torch.float64, device="cuda")


def _loopback_solve(hp, util, world, impl, tol=1e-8, nu0=None):
    lib = _lib.load()
    n = hp.n_tokens
    spec = util.spec(n)
    stores = [cf.PoolStore(hp, rank=r, world=world) for r in range(world)]
    assert all(len(s.buckets) == 1 and s.buckets[0].blocked for s in stores)
    # receive areas [3 slots][world sources][stride] cells of 16 B, zeroed; the device arrays of their base pointers
    acc_areas = [torch.zeros(3 * world * (n + 1) * 2, **F64) for _ in range(world)]
    vec_areas = [torch.zeros(3 * world * (n + 2) * 2, **F64) for _ in range(world)]
    acc_ptrs = torch.tensor([t.data_ptr() for t in acc_areas], dtype=torch.int64, device="cuda")
    vec_ptrs = torch.tensor([t.data_ptr() for t in vec_areas], dtype=torch.int64, device="cuda")
    c = torch.as_tensor(np.asarray(spec.c, float), **F64)
    a = torch.as_tensor(np.asarray(spec.a, float), **F64)
    eq = torch.as_tensor(np.asarray(spec.eq, np.uint8), device="cuda")
    pinned = torch.as_tensor(np.asarray(spec.pinned, np.uint8), device="cuda")
    start = default_nu0(spec) if nu0 is None else np.asarray(nu0, float)
    nus = [torch.as_tensor(start, **F64).clone() for _ in range(world)]
    psis = [torch.empty(n, **F64) for _ in range(world)]
    work_bytes = lib.cfmm_persist_solve_work_bytes if impl == "persist" else lib.cfmm_blocked_solve_work_bytes
    entry = lib.cfmm_persist_solve if impl == "persist" else lib.cfmm_blocked_solve_peer
    works = [torch.empty(int(work_bytes(C.byref(s.buckets[0].c_blocked), n)), dtype=torch.uint8, device="cuda") for s in stores]
    streams = [torch.cuda.Stream() for _ in range(world)]
    scale = max(float(np.abs(spec.c).max()), 1.0)
    results = [None] * world
    torch.cuda.synchronize()

    def rank_main(r):
        prm = _lib.SolveParams(float(tol), 1e-12 * scale, 100, 200)
        res = _lib.SolveResult()
        pc = _lib.PeerCtx(int(acc_ptrs.data_ptr()), int(vec_ptrs.data_ptr()), r, world, 0, 0)
        rc = entry(C.byref(stores[r].buckets[0].c_blocked), n, c.data_ptr(), a.data_ptr(), eq.data_ptr(), pinned.data_ptr(),
                   nus[r].data_ptr(), psis[r].data_ptr(), works[r].data_ptr(), C.byref(prm), C.byref(res), C.byref(pc),
                   C.c_void_p(streams[r].cuda_stream))
        results[r] = (rc, res.status, res.primal_value, res.dual_value, res.gap, res.iters, res.evals, res.hvps, pc.seq_acc, pc.seq_vec)

    threads = [threading.Thread(target=rank_main, args=(r,), daemon=True) for r in range(world)]   # daemon: a stuck rank fails the test, not the exit
    assert lib.cfmm_set_persist_cooperative(0) == 0
    try:
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=30)
        torch.cuda.synchronize()
    finally:
        lib.cfmm_set_persist_cooperative(1)
    assert all(not t.is_alive() for t in threads), "a loopback rank did not return"
    return results, [x.cpu().numpy() for x in nus], [x.cpu().numpy() for x in psis]


@pytest.mark.timeout(120)
@pytest.mark.parametrize("impl", ["persist", "hostloop"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_solvers_over_loopback_ranks_match_the_single_store_solve(world, impl):
    # 29k pools = 33 tiles in total: at world 8 every rank runs 4-5 CTAs, all co-resident; 1000 tokens = 63 slices
    hp, s = H.cp_host_pools(29_000, 1000, seed=4)
    for util, nu0 in ((cf.Arbitrage(s["prices"]), None), (cf.Liquidate(0, np.where(np.arange(1000) % 97 == 5, 3.0, 0.0)), s["prices"] / s["prices"][0])):
        one = cf.solve_pools(hp, util, nu0=nu0, tol=1e-8, native=impl, want_trades=False)
        assert one.status == "optimal"
        results, nus, psis = _loopback_solve(hp, util, world, impl, nu0=nu0)
        for r, (rc, status, primal, dual, gap, iters, evals, hvps, sa, sv) in enumerate(results):
            assert rc == 0 and status == 0, (world, impl, r, rc, status)
            assert abs(primal - one.value) <= 1e-8 * abs(one.value) + 1e-9 * abs(one.dual_value)
        # every rank reduced the same vectors in the same (rank) order: identical decisions, identical bits
        for r in range(1, world):
            assert results[r][1:] == results[0][1:], (results[0], results[r])
            assert np.array_equal(nus[r], nus[0]) and np.array_equal(psis[r], psis[0])
        assert np.abs(psis[0] - one.psi).max() <= 1e-6 * max(np.abs(one.psi).max(), 1.0)
        # the exchanges of a solve: one per evaluation on the [psi | arb] channel; one per Hessian product and diagonal
        evals, hvps, iters = results[0][6], results[0][7], results[0][5]
        assert results[0][8] == evals and results[0][9] >= hvps
