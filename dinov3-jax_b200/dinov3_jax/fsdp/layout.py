"""Shard layout of the flat parameter buffers for FSDP over the "dp" axis (pure Python: testable on CPU).

The reference shards every parameter leaf of a wrapped unit on its largest divisible axis (dinov3_jax/fsdp/utils.py:
19-53) and all-gathers leaf by leaf (:56-84).  Here a unit (patch-embed + tokens, each transformer block, the final
norm, each head — models/vision_transformer.py:93,137; train/ssl_meta_arch.py:77-78,122-123) owns two contiguous
ranges of its module's flat buffer (matrices, vectors); rank r holds the r-th 1/world slice of each range.  After a
tiled all-gather the full range is laid out exactly like the single-GPU buffer, so kernels never see the sharding.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class UnitRange:
    name: str
    mat: tuple          # [start, end) in the module's flat buffer (matrix region)
    vec: tuple          # [start, end) in the module's flat buffer (vector region)
    tensors: tuple      # tensor names of the unit


def unit_of(module: str, tensor: str) -> str:
    if module != "backbone":
        return "head"
    if tensor.startswith("blocks_"):
        return tensor.split("/")[0]
    if tensor.startswith("norm/"):
        return "norm"
    return "embed"          # patch_embed/*, cls_token, mask_token


class ShardLayout:
    """offsets/sizes: name -> element offset / padded size in the module flat buffer (matrices first, then vectors);
    n_mat: size of the matrix region; n: total size."""

    def __init__(self, module: str, names_in_order, offsets: dict, padded: dict, kinds: dict, n_mat: int, n: int,
                 world: int):
        assert world >= 1
        self.module, self.world, self.n_mat, self.n = module, world, n_mat, n
        self.offsets, self.padded, self.kinds = offsets, padded, kinds
        order = []
        groups = {}
        for nm in names_in_order:
            u = unit_of(module, nm)
            if u not in groups:
                groups[u] = []
                order.append(u)
            groups[u].append(nm)
        self.units = []
        for u in order:
            mats = [nm for nm in groups[u] if kinds[nm] == "mat"]
            vecs = [nm for nm in groups[u] if kinds[nm] == "vec"]
            rng = lambda lst, empty_at: ((min(offsets[k] for k in lst), max(offsets[k] + padded[k] for k in lst))
                                         if lst else (empty_at, empty_at))
            self.units.append(UnitRange(u, rng(mats, 0), rng(vecs, n_mat), tuple(groups[u])))
        # every range must split into `world` slices of a multiple of 8 elements (16 B in bf16, float4-able in fp32)
        for u in self.units:
            for a, b in (u.mat, u.vec):
                assert (b - a) % (8 * world) == 0, f"unit {u.name}: range {a}:{b} not divisible by 8*world={8 * world}"
        # contiguity: consecutive units tile each region without gaps
        for region, lo, hi in (("mat", 0, n_mat), ("vec", n_mat, n)):
            cur = lo
            for u in self.units:
                a, b = getattr(u, region)
                if b > a:
                    assert a == cur, (module, u.name, region, a, cur)
                    cur = b
            assert cur == hi, (module, region, cur, hi)
        # shard buffer layout: [mat slices of all units | vec slices of all units]
        self.shard_mat_off, self.shard_vec_off = {}, {}
        off = 0
        for u in self.units:
            self.shard_mat_off[u.name] = off
            off += (u.mat[1] - u.mat[0]) // world
        self.n_mat_shard = off
        for u in self.units:
            self.shard_vec_off[u.name] = off
            off += (u.vec[1] - u.vec[0]) // world
        self.n_shard = off
        assert self.n_shard * world == n

    def slice_of(self, unit: UnitRange, region: str, rank: int):
        """[start, end) in the FULL flat buffer held by `rank` for this unit's region."""
        a, b = getattr(unit, region)
        s = (b - a) // self.world
        return a + rank * s, a + (rank + 1) * s

    def shard_range(self, unit: UnitRange, region: str):
        """[start, end) in the per-rank SHARD buffer for this unit's region."""
        a, b = getattr(unit, region)
        s = (b - a) // self.world
        off = (self.shard_mat_off if region == "mat" else self.shard_vec_off)[unit.name]
        return off, off + s

    def full_to_shard_index(self, rank: int) -> np.ndarray:
        """int64 [n_shard]: full-buffer index of every element of rank's shard buffer (tests / import / export)."""
        idx = np.empty(self.n_shard, dtype=np.int64)
        for region in ("mat", "vec"):
            for u in self.units:
                fa, fb = self.slice_of(u, region, rank)
                sa, sb = self.shard_range(u, region)
                idx[sa:sb] = np.arange(fa, fb)
        return idx

    def shard_segments(self, rank: int, multipliers: dict):
        """Optimiser segment table in shard coordinates: (start, lr_mult, wd_mult, is_last) per maximal run of one
        tensor inside rank's shard, sorted by start.  `multipliers`: tensor name -> (lr_mult, wd_mult, is_last)."""
        segs = []
        for region in ("mat", "vec"):
            for u in self.units:
                fa, fb = self.slice_of(u, region, rank)
                sa, _ = self.shard_range(u, region)
                for nm in u.tensors:
                    if self.kinds[nm] != region:
                        continue
                    ta, tb = self.offsets[nm], self.offsets[nm] + self.padded[nm]
                    lo, hi = max(ta, fa), min(tb, fb)
                    if lo < hi:
                        segs.append((sa + (lo - fa),) + tuple(multipliers[nm]))
        segs.sort(key=lambda t: t[0])
        assert segs and segs[0][0] == 0
        return segs
