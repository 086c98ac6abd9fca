"""Destination-sharded multi-GPU regridder (one process per GPU, ``torch.distributed``).

SURVEY.md section 8(e): destination cells are split into contiguous field-index blocks, one per
rank; each rank builds the row block of ``A`` for its destination cells against the *replicated*
source grid -- no exchange during the build.  ``transpose(R)`` is served by a second local
build with the roles swapped (row block of ``A^T`` for the rank's *source* cells), so neither
direction needs a reduction.  Collectives (NCCL over NVLink on GPUs, gloo in the CPU tests) are
used only to broadcast the input field and to all-gather the output field / the area vectors.

The reference has no distributed path at all; the single-process semantics
this reproduces are ``Regridder`` / ``regrid!`` / ``transpose`` (src/regridder/regridder.jl:125-163,
49-50; src/regridder/regrid.jl:63-118).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .grids import Grid


def block_bounds(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, near-equal blocks of ``range(n)``; the first ``n % world`` blocks get one more."""
    q, r = divmod(n, world)
    out, lo = [], 0
    for k in range(world):
        hi = lo + q + (1 if k < r else 0)
        out.append((lo, hi))
        lo = hi
    return out


class _LocalB200:
    """Row-block operator backed by the CUDA engine: y_shard = (A_block x) ./ areas_block."""

    def __init__(self, rows_grid: Grid, cols_grid: Grid, stream: Optional[int] = None, **kw):
        from .regridder import Regridder
        self.R = Regridder(rows_grid, cols_grid, build_transpose=False, stream=stream, **kw)
        self.areas = torch.from_numpy(self.R.dst_areas)
        self.nnz = self.R.intersections.nnz
        self.stats = self.R.intersections.stats()

    def apply(self, out: torch.Tensor, x: torch.Tensor, normalize: bool = True):
        from .regridder import regrid_
        regrid_(out, self.R, x, normalize=normalize, asynchronous=out.is_cuda)


class ShardedRegridder:
    """``Regridder(dst, src)`` sharded over the ranks of ``group`` by destination cells.

    ``local_factory(rows_grid, cols_grid) -> op`` builds a rank-local row-block operator with
    ``op.apply(out, x, normalize)``, ``op.areas`` (torch, length = rows) and ``op.nnz``; the
    default is the CUDA engine.  (The CPU tests inject a numpy/scipy factory to exercise the
    sharding and the collectives under gloo.)"""

    def __init__(self, dst: Grid, src: Grid, group=None, local_factory: Optional[Callable] = None,
                 device: Optional[torch.device] = None, build_transpose: bool = True):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_dst, self.n_src = dst.ncells, src.ncells
        self.device = device if device is not None else torch.device("cpu")
        factory = local_factory or _LocalB200
        self.dst_bounds = block_bounds(self.n_dst, self.world)
        self.src_bounds = block_bounds(self.n_src, self.world)
        lo, hi = self.dst_bounds[self.rank]
        self.fwd = factory(dst.slice(lo, hi), src)
        self.bwd = None
        if build_transpose:
            slo, shi = self.src_bounds[self.rank]
            self.bwd = factory(src.slice(slo, shi), dst)
        # R.dst_areas / R.src_areas: all-gather of the per-shard geometric areas
        self.dst_areas = self._all_gather_blocks(self.fwd.areas.to(self.device), self.dst_bounds)
        self.src_areas = self._all_gather_blocks(self.bwd.areas.to(self.device), self.src_bounds) \
            if self.bwd is not None else None
        nnz = torch.tensor([self.fwd.nnz], dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.all_reduce(nnz, group=self.group)
        self.nnz = int(nnz.item())

    @property
    def shape(self):
        return (self.n_dst, self.n_src)

    # -- collectives -------------------------------------------------------------------------
    def _all_gather_blocks(self, shard: torch.Tensor, bounds) -> torch.Tensor:
        """Concatenate per-rank blocks (sizes differ by at most one) into the full vector(s).
        ``shard`` has the block rows in its first dimension."""
        n = bounds[-1][1]
        if self.world == 1:
            return shard.clone()
        width = max(hi - lo for lo, hi in bounds)
        pad = torch.zeros((width,) + tuple(shard.shape[1:]), dtype=shard.dtype, device=shard.device)
        pad[: shard.shape[0]] = shard
        full = torch.empty((self.world * width,) + tuple(shard.shape[1:]), dtype=shard.dtype, device=shard.device)
        dist.all_gather_into_tensor(full, pad, group=self.group)
        if all(hi - lo == width for lo, hi in bounds):
            return full[:n]
        return torch.cat([full[k * width: k * width + (hi - lo)] for k, (lo, hi) in enumerate(bounds)])

    def _broadcast(self, x: Optional[torch.Tensor], n: int, trailing=(), root: int = 0) -> torch.Tensor:
        if self.world == 1:
            return x
        if x is None or self.rank != root:
            x = torch.empty((n,) + tuple(trailing), dtype=torch.float64, device=self.device)
        dist.broadcast(x, src=dist.get_global_rank(self.group, root) if self.group is not None else root,
                       group=self.group)
        return x

    # -- regrid! -------------------------------------------------------------------------------
    def regrid(self, src_field: Optional[torch.Tensor], transpose: bool = False, normalize: bool = True,
               broadcast: bool = True, gather: bool = True, trailing=()) -> torch.Tensor:
        """``regrid!``: the input field (length n_src, or (n_src, K) level-fastest) lives on rank 0
        and is broadcast (unless ``broadcast=False``: every rank already holds it); each rank
        computes its block; the blocks are all-gathered (unless ``gather=False``: returns the
        local block)."""
        op = self.bwd if transpose else self.fwd
        if op is None:
            raise ValueError("built with build_transpose=False")
        n_in = self.n_dst if transpose else self.n_src
        bounds = self.src_bounds if transpose else self.dst_bounds
        x = self._broadcast(src_field, n_in, trailing) if broadcast else src_field
        lo, hi = bounds[self.rank]
        out = torch.zeros((hi - lo,) + tuple(x.shape[1:]), dtype=torch.float64, device=x.device)
        if hi > lo:
            op.apply(out, x, normalize)
        return self._all_gather_blocks(out, bounds) if gather else out
