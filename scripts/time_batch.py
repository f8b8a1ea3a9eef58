"""Times cfmm_batch_solve (one whole solve per thread): the two-asset.py sweep (50 problems) and large quote batches."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import cfmm_routing_code_b200 as cf
from cfmm_routing_code_b200 import batch as B, instances as I


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    out = {}
    d = I.two_asset_instance()
    hp = cf.HostPools.from_lists(d["n_tokens"], d["local_indices"], d["reserves"], d["fees"], d["kinds"], d["weights"])
    store = cf.CsrStore(hp)
    for nb in (50, 4096, 131072, 1048576):
        ts = np.linspace(0.0, 50.0, nb)
        us = [cf.Swap(0, 2, t) for t in ts] if nb <= 4096 else None
        if us is not None:
            c, a, fl, nu0 = B.pack_utilities(us, hp.n_tokens)
        else:
            c = np.zeros((nb, 3)); c[:, 2] = 1.0
            a = np.zeros((nb, 3)); a[:, 0] = ts
            fl = np.zeros((nb, 3), np.uint8); nu0 = np.ones((nb, 3))
        cd, ad, fd = (torch.as_tensor(x, device="cuda") for x in (c, a, fl))
        nu_init = torch.as_tensor(nu0, device="cuda")
        res = {}

        def run():
            nu = nu_init.clone()
            res["r"] = cf.solve_batch_device(store, cd, ad, fd, nu, tol=1e-9, want_trades=False)
        ms = timed(run)
        st = res["r"][1].cpu().numpy()
        out[f"two_asset_sweep_B{nb}"] = dict(ms=ms, problems_per_s=nb / ms * 1e3, optimal=int((st[:, 7] == 0).sum()),
                                             evals_mean=float(st[:, 6].mean()), evals_max=float(st[:, 6].max()),
                                             u_first=float(st[0, 0]), u_last=float(st[-1, 0]))
    # end to end through the public call, host lists in, Result objects out
    us = [cf.Swap(0, 2, t) for t in d["amounts"]]
    args = (d["local_indices"], d["reserves"], d["fees"], d["kinds"], d["weights"], us)
    for batched in (True, False):
        cf.solve_sweep(*args, tol=1e-9, batched=batched)
        t0 = time.perf_counter(); rs = cf.solve_sweep(*args, tol=1e-9, batched=batched); dt = time.perf_counter() - t0
        out[f"solve_sweep_batched_{batched}"] = dict(wall_ms=dt * 1e3, u50=rs[-1].value)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
