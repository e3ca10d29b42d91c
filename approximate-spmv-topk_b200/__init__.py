"""approximate-spmv-topk_b200 -- B200-native fused Top-K SpMV engine (Python face).

The product is libtopkspmv.so (hand-written sm_100a CUDA behind the C ABI of include/topkspmv.h) plus
the C++ host surface under host/.  This package is the thin Python mirror the tests, bench.py and the
sweep driver use: ctypes bindings (capi), the `SpMV` functor with the reference's four verbs (spmv),
the synthetic matrix generator (create_matrices), row-shard planning (sharding) and the one-process-per-GPU
candidate exchange (distributed).

The directory name carries a hyphen, so import it through the repo-root shim:

    from _pkg import pkg; tks = pkg()          # registers it as `approximate_spmv_topk_b200`
"""
from . import accuracy, capi, create_matrices, distributed, sharding, spmv  # noqa: F401
from .spmv import SpMV, SpMVFixed  # noqa: F401

from .distributed import ShardedSpMV, ShardedSpMVFixed  # noqa: F401

__all__ = ["accuracy", "capi", "create_matrices", "distributed", "sharding", "spmv", "SpMV", "SpMVFixed", "ShardedSpMV", "ShardedSpMVFixed"]
