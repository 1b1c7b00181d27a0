"""`ac_compile_parallelize` (dinov3_jax/fsdp/ac_compile_parallelize.py:20-44): eager sharding of the parameter tree over
the "dp" axis — every leaf with two or more dimensions is split on its largest axis that the axis size divides
(ties resolved like `np.argsort(shape)[::-1]`), 1-D leaves stay replicated; `min_shard_size` is accepted and, as in
the reference, not consulted.  The reference places the shards with `jax.device_put(NamedSharding)`; here each rank
keeps its slice as a `Partitioned` box (fsdp/utils.py).  The training engine does not go through this function (it
shards by FSDP unit, fsdp/layout.py) — it exists for callers of the reference API (`prepare_for_distributed_training`).
"""
from __future__ import annotations

import numpy as np

from .utils import Partitioned, _axis_index, _axis_size, _map


def ac_compile_parallelize(trained_model, inference_only_models=None, config=None, min_shard_size=2 ** 12):
    idx, size = _axis_index("dp"), _axis_size("dp")

    def shard(p):
        if isinstance(p, Partitioned) or p.dim() <= 1:
            return p
        for i in np.argsort(p.shape)[::-1]:
            if p.shape[i] % size == 0:
                s = p.shape[i] // size
                return Partitioned(p.narrow(int(i), idx * s, s).contiguous(), int(i), "dp")
        return p                                     # no divisible axis: replicated (:33-34)
    return _map(shard, trained_model)
