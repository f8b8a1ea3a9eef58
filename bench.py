#!/usr/bin/env python
"""bench.py -- pools x dual-evaluations / second on BASELINE.json configs[4] (1M constant-product pools,
4096 tokens), plus wall-clock to 1e-6 relative gap, the HBM roofline of the dominant kernel, and the CPU
baseline timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--scaling weak|strong]

A "step" is one dual evaluation: the per-pool optimal-arbitrage kernel over this rank's pools, accumulating
psi(nu) and the dual value (+ one all-reduce of the (n_tokens+1)-vector when N > 1).  Steps rotate over 8
independent pool instances per GPU (8 x 32 MB = 256 MB > the 126 MB L2), so every step streams its pools from
HBM.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_POOLS = 1_000_000
N_TOKENS = 4096
N_INSTANCES = 8
METRIC = "pools x dual-evaluations / second (1M constant-product pools, 4096 tokens); time to 1e-6 rel-gap reported beside it"
UNIT = "pool-evals/s"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML; nvidia-smi as a fallback)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.sm, self.reasons, self.sm_max = [], set(), None
        self.stop = threading.Event()
        self.index = index
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None
        self.th = threading.Thread(target=self._run, daemon=True)

    @staticmethod
    def _physical_index(i):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[i])
            except Exception:
                return i
        return i

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("hw_thermal_slowdown", 0x40),
                          ("sw_thermal_slowdown", 0x20)):
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                              "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
        r = [x.strip() for x in out.strip().split(",")]
        if len(r) >= 6:
            self.sm.append(float(r[0])); self.sm_max = float(r[1])
            for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
                if r[2 + i] == "Active":
                    self.reasons.add(name)

    def _run(self):
        while not self.stop.is_set():
            try:
                self._sample_nvml() if self.nvml else self._sample_smi()
            except Exception:
                pass
            self.stop.wait(0.02)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml" if self.nvml else "nvidia-smi"}


# --------------------------------------------------------------------------------------------------
# CPU legs (the ONLY places this file executes oracle/)
# --------------------------------------------------------------------------------------------------
def cpu_eval_throughput(seconds_budget=12.0):
    """oracle dual evaluations of the full 1M-pool instance on ALL host cores (oracle/cfmm_oracle_c.c, pthreads):
    pool-evals/s, number of evaluations, seconds, threads."""
    from cfmm_routing_code_b200 import instances as I
    from oracle import c_oracle as CO
    s = I.synth_const_product(M_POOLS, N_TOKENS, seed=3)
    idx = np.ascontiguousarray(s["idx"], np.int32); R = np.ascontiguousarray(s["reserves"]); g = s["gamma"]
    nu = s["prices"] * np.exp(0.01 * np.random.default_rng(0).standard_normal(N_TOKENS))
    CO.autotune_threads(idx, R, g, N_TOKENS, nu)      # best thread count for this host (also warms up)
    t0 = time.perf_counter(); k = 0
    while True:
        CO.eval_pairs(idx, R, g, N_TOKENS, nu * (1 + 1e-3 * k)); k += 1
        dt = time.perf_counter() - t0
        if dt > seconds_budget or k >= 2000:
            break
    return M_POOLS * k / dt, k, dt, CO.num_threads()


def cpu_full_solve():
    """the same dual algorithm end to end on the host cores: solver.py's loop over the C/pthreads oracle evaluator"""
    import cfmm_routing_code_b200 as cf
    from cfmm_routing_code_b200 import instances as I
    from cfmm_routing_code_b200.solver import solve_dual
    from oracle import c_oracle as CO
    s = I.synth_const_product(M_POOLS, N_TOKENS, seed=3)
    nu = s["prices"].copy()
    CO.autotune_threads(np.ascontiguousarray(s["idx"], np.int32), np.ascontiguousarray(s["reserves"]), s["gamma"],
                        N_TOKENS, nu)
    t0 = time.perf_counter()
    ev = CO.CpuPairsEvaluator(N_TOKENS, s["idx"], s["reserves"], s["gamma"])
    r = solve_dual(ev, cf.Arbitrage(s["prices"]).spec(N_TOKENS), tol=1e-6, linear_solver="cg")
    wall = time.perf_counter() - t0
    return {"value": M_POOLS * r.evals / wall, "unit": UNIT, "wall_s": wall, "evals": r.evals, "hvps": r.hvps,
            "status": r.status, "gap": r.gap, "threads": CO.num_threads()}


def run_reference(args):
    """The reference's path on the host cores.  cvxpy (the reference's solver) is probed at run time; it is
    not in this image, so the oracle port (same dual evaluation, numpy, 1 thread) stands in -- kind 'port'."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    try:
        import cvxpy  # noqa: F401
        have_cvxpy = True
    except Exception:
        have_cvxpy = False
    vals = []
    per_step_budget = max(1.0, min(15.0, 90.0 / max(args.steps + args.warmup, 1)))
    cores = 1
    for i in range(args.warmup + args.steps):
        v, k, dt, cores = cpu_eval_throughput(per_step_budget)
        if i >= args.warmup:
            vals.append((v, k, dt))
    value = float(np.mean([v for v, _, _ in vals]))
    sample = (f"{vals[0][1]} oracle dual evaluations of the full 1M-pool/4096-token instance per step "
              f"(C restatement oracle/cfmm_oracle_c.c, {cores} pthreads, fp64)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([dt / k for _, k, dt in vals])),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args.gpus, args.scaling),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "reference solver (cvxpy) " + ("present but not used for this metric" if have_cvxpy
                                                               else "unavailable in image")},
        "gpu_launches": 0,
    }
    solve = cpu_full_solve()
    line["e2e"] = {"value": solve["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                   "what": "one full solve to a 1e-6 certificate on the host cores (same algorithm: solver.py over the "
                           "C/pthreads oracle evaluator); value = pools x dual evaluations / wall",
                   "wall_s": solve["wall_s"], "evals": solve["evals"], "hvps": solve["hvps"], "status": solve["status"]}
    line["time_to_1e-6_gap"] = {"seconds": solve["wall_s"], "rel_gap": abs(solve["gap"]), "tol": 1e-6}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, scaling):
    per = M_POOLS if scaling == "weak" else M_POOLS // n_gpus
    return {"workload": "BASELINE.json configs[4]: 1M constant-product pools, 4096 tokens, Arbitrage(c=p), seed 3",
            "pools_per_gpu": per, "n_tokens": N_TOKENS, "parallelism": f"pool-shard x{n_gpus}",
            "l2": f"rotating {N_INSTANCES} pool instances per GPU ({N_INSTANCES * per * 32 // 2**20} MiB) > 126 MB L2",
            "collective": "none" if n_gpus == 1 else "one all-reduce of n_tokens+1 f64 per step"}


# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import cfmm_routing_code_b200 as cf
    from cfmm_routing_code_b200 import instances as I

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU leg")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    per = M_POOLS if args.scaling == "weak" else M_POOLS // world
    f64 = dict(dtype=torch.float64, device=dev)

    stores, nus = [], []
    for k in range(N_INSTANCES):
        # weak: every rank owns `per` pools of its own (seeded by rank); strong: rank's slice of the seed-3 instance
        if args.scaling == "weak":
            s = I.synth_const_product(per, N_TOKENS, seed=3 + 100 * k + 7919 * rank)
        else:
            s = I.synth_const_product(M_POOLS, N_TOKENS, seed=3 + 100 * k)
            sl = slice(rank * per, (rank + 1) * per)
            s = dict(s, idx=s["idx"][sl], reserves=s["reserves"][sl], gamma=s["gamma"][sl])
        hp = cf.HostPools.from_pairs(N_TOKENS, s["idx"], s["reserves"], s["gamma"])
        stores.append(cf.PoolStore(hp, device=dev, validate=False))
        if world > 1 and args.collective == "peer":
            try:
                stores[-1].enable_peer_allreduce()
            except Exception as e:                       # symmetric memory unavailable: NCCL does the all-reduce
                if rank == 0:
                    print(f"peer all-reduce unavailable ({type(e).__name__}: {e}); falling back to NCCL", file=sys.stderr)
                args.collective = "nccl"
        p = I.synth_const_product(8, N_TOKENS, seed=3)["prices"]        # same token prices on every rank
        nus.append(torch.as_tensor(p * np.exp(0.01 * np.random.default_rng(k).standard_normal(N_TOKENS)), **f64))
    lib = stores[0].lib

    peer = world > 1 and args.collective == "peer"

    def step(i):
        acc = stores[i % N_INSTANCES].evaluate(nus[i % N_INSTANCES])      # peer mode: already all-reduced
        if world > 1 and not peer:
            dist.all_reduce(acc)
        return acc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    # N = 1: the K steps are captured once into CUDA graphs of CHUNK steps each and replayed, so the timed
    # region holds kernel work only (no Python / ctypes launch overhead between the ~10 us kernels).
    CHUNK = 64
    use_graph = not args.no_graph and args.steps >= CHUNK and (world == 1 or peer or args.graph_nccl)
    steps = (args.steps // CHUNK) * CHUNK if use_graph else args.steps
    graph = None
    if use_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(N_INSTANCES):
                step(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(CHUNK):
                step(i)
        graph.replay()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib.cfmm_reset_launch_count()
    with ClockSampler(local) as clocks:
        barrier()
        torch.cuda.profiler.start()
        e0.record()
        if use_graph:
            for _ in range(steps // CHUNK):
                graph.replay()
        else:
            for i in range(steps):
                step(i)
        e1.record()
        barrier()
        torch.cuda.profiler.stop()
    launches = int(lib.cfmm_launch_count()) if not use_graph else steps * sum(len(st.buckets) for st in stores[:1])
    ms = torch.tensor([e0.elapsed_time(e1)], **f64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms)
    ms_per_step = total_ms / steps
    value = n_gpus * per / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (k_eval_pair<PRODUCT>): algorithmic bytes / avg launch duration
    alg_bytes = stores[0].algorithmic_bytes_per_eval()
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tfile):
        try:
            traffic = json.load(open(tfile)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "k_blocked<eval> (csrc/cfmm_blocked.cu)",
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "note": "duration = timed region / steps: one kernel node per step incl. launch gaps" +
                        (", all-reduce" if world > 1 else "") + (", CUDA-graph replay" if use_graph else "")}

    # ---- e2e: the public API on HOST buffers: upload pools, solve to 1e-6, read psi/nu back
    e2e = None
    time_to_gap = None
    if (rank == 0 or world > 1) and not args.no_e2e:
        s = I.synth_const_product(M_POOLS, N_TOKENS, seed=3)
        hp = cf.HostPools.from_pairs(N_TOKENS, s["idx"], s["reserves"], s["gamma"]).pin_memory()
        runs = []
        for rep in range(3):
            barrier()
            t0 = time.perf_counter()
            if world > 1:
                st_e = cf.PoolStore(hp, device=dev, rank=rank, world=world, validate=False)
                if peer:
                    st_e.enable_peer_allreduce()
                r = cf.solve_pools(hp, cf.Arbitrage(s["prices"]), tol=1e-6, want_trades=False, store=st_e)
            else:
                r = cf.solve_pools(hp, cf.Arbitrage(s["prices"]), tol=1e-6, want_trades=False, device=dev)
            torch.cuda.synchronize()
            runs.append((time.perf_counter() - t0, r))
        wall, r = min(runs, key=lambda x: x[0])
        if world > 1:
            w = torch.tensor([wall], **f64); dist.all_reduce(w, op=dist.ReduceOp.MAX); wall = float(w)
        h2d = hp.reserves.nbytes + hp.tok_idx.nbytes + hp.gamma.nbytes + 8 * N_TOKENS
        e2e = {"value": M_POOLS * r.evals / wall, "unit": UNIT, "h2d_bytes_per_step": int(h2d / max(world, 1)),
               "d2h_bytes_per_step": 16 * N_TOKENS + 64,
               "what": "cf.solve_pools(pinned host numpy pools, Arbitrage(p), tol=1e-6): upload + layout build + solve + psi/nu read-back; "
                       "value = pools x dual evaluations / wall",
               "wall_s": wall, "evals": r.evals, "hvps": r.hvps, "iters": r.iters, "status": r.status,
               "gap": r.gap, "primal_infeas": r.primal_infeas}
        time_to_gap = {"seconds_incl_upload": wall, "seconds_solver_only": r.wall_s, "rel_gap": abs(r.gap),
                       "primal_infeas": r.primal_infeas, "tol": 1e-6}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, k, dt, cores = cpu_eval_throughput(12.0)
        solve = cpu_full_solve()
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "time_to_1e-6_gap_s": solve["wall_s"],
               "e2e_value": solve["value"],
               "sample": f"{k} oracle dual evaluations (C restatement, {cores} pthreads, fp64) of the same 1M-pool "
                         f"instance in {dt:.1f}s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(n_gpus, args.scaling),
                           **({"collective_impl": ("cfmm_allreduce_ll: 16-byte {value, seq} pushes over NVLink peer memory, PDL-chained between the evaluation kernels" if peer
                                                   else "NCCL all_reduce")} if world > 1 else {})),
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": e2e, "time_to_1e-6_gap": time_to_gap, "gpu_launches": launches, "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        sys.stdout.flush()
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)          # captured graphs + symmetric memory: skip the (hang-prone) communicator teardown


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer solve leg (profiling runs)")
    ap.add_argument("--no-graph", action="store_true", help="launch every step from Python instead of CUDA-graph replay")
    ap.add_argument("--graph-nccl", action="store_true", help="N>1: capture the all-reduce into the step graphs too")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="N>1: fused one-shot NVLink all-reduce (cfmm_allreduce_oneshot) or NCCL")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 20:
            args.steps = 10         # each reference step is seconds of CPU work
            args.warmup = min(args.warmup, 1)
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
