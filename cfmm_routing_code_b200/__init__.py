"""cfmm_routing_code_b200 -- B200-native optimal routing for constant-function market makers.

A drop-in for the `prob.solve()` call of angeris/cfmm-routing-code's scripts (arbitrage.py:82,
liquidation.py:85, two-asset.py:91): same problem literals in, same (value, psi, deltas, lambdas) out,
computed by dual decomposition with hand-written sm_100a kernels (libcfmm_b200.so, C ABI in
include/cfmm_b200.h).  There is no CPU path.
"""
from .api import Arbitrage, Liquidate, Swap, LinearUtility, Result, solve, solve_pools, solve_sweep   # noqa: F401
from .pools import HostPools, PoolStore                                    # noqa: F401
from .batch import CsrStore, solve_batch, solve_batch_device, solve_many               # noqa: F401
from .solver import DualSpec, solve_dual                                   # noqa: F401
from ._lib import CfmmError                                                # noqa: F401

__version__ = "0.1.0"
