"""Shared builders for tests (not a test module)."""
import numpy as np

import cfmm_routing_code_b200 as cf
from cfmm_routing_code_b200 import instances as I
from oracle import cfmm_oracle as O


def host_pools(d):
    return cf.HostPools.from_lists(d["n_tokens"], d["local_indices"], d["reserves"], d["fees"], d["kinds"],
                                   d["weights"])


def oracle_pools(hp):
    return O.Pools(hp.n_tokens, hp.pool_ptr, hp.tok_idx, hp.reserves, hp.weights, hp.gamma, hp.kind)


def cp_host_pools(m, n, seed):
    s = I.synth_const_product(m, n, seed)
    return cf.HostPools.from_pairs(n, s["idx"], s["reserves"], s["gamma"]), s


def mixed_host_pools(m, n, seed):
    s = I.synth_mixed(m, n, seed)
    hp = cf.HostPools(n, s["pool_ptr"], s["tok_idx"], s["reserves"], s["weights"], s["gamma"], s["kind"])
    return hp, s


def random_prices(prices, seed, spread=0.05):
    rng = np.random.default_rng(seed)
    return prices * np.exp(spread * rng.standard_normal(len(prices)))
