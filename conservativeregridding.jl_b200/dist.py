"""Destination-sharded multi-GPU regridder (one process per GPU, ``torch.distributed``).

SURVEY.md section 8(e): destination cells are split into contiguous field-index blocks, one per
rank; each rank builds the row block ``A_r`` of ``A`` for its destination cells against the
*replicated* source grid -- ONE local build, no exchange during the build.

* forward ``regrid!``: the source field is broadcast (NCCL), every rank computes its block
  ``y_r = (A_r x) ./ a_dst_r`` and the blocks are all-gathered;
* ``transpose(R)``: ``A^T y = sum_r A_r^T y_r`` -- every rank applies the transpose of its own block
  (the CSR(A_r^T) that the local assembly produces anyway) to its slice of the destination field
  and the partial source vectors are summed with one all-reduce over NVLink; the division by the
  (replicated, geometric) source areas follows the reduction.
* ``R.dst_areas``: all-gather of the per-block areas; ``R.src_areas``: computed by every rank.
* ``normalize=True``: every block is scaled by the all-reduced (max) ``maximum(A_r)``.

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
    """Row-block operator backed by the CUDA engine (one ``crg_regridder`` handle)."""

    def __init__(self, rows_grid: Grid, cols_grid: Grid, stream: Optional[int] = None, **kw):
        from .regridder import Regridder, transpose
        self.R = Regridder(rows_grid, cols_grid, build_transpose=True, stream=stream, **kw)
        self.RT = transpose(self.R)
        self.nnz = self.R.intersections.nnz
        self.stats = self.R.intersections.stats()

    def _areas(self, which: str, device):
        n = self.R.shape[0] if which == "dst" else self.R.shape[1]
        if device is not None and torch.device(device).type == "cuda":      # device to device, no host round trip
            from .regridder import areas_to
            t = torch.empty(n, dtype=torch.float64, device=device)
            areas_to(self.R, **{f"{which}_areas_out": t})
            return t
        return torch.from_numpy(self.R.dst_areas if which == "dst" else self.R.src_areas)

    def areas(self, device=None):          # geometric areas of this rank's destination block
        return self._areas("dst", device)

    def src_areas(self, device=None):      # geometric areas of the (replicated) source grid
        return self._areas("src", device)

    def maximum(self) -> float:           # maximum(A_r); 0 for an empty block
        return self.R.intersections.maximum()

    def scale(self, divisor: float):      # A_r, A_r^T and both area vectors ./= divisor
        from .regridder import scale_
        scale_(self.R, divisor)

    def apply(self, out: torch.Tensor, x: torch.Tensor, normalize: bool = True):
        """out = (A_r x) ./ a_dst_r"""
        from .regridder import regrid_
        regrid_(out, self.R, x, normalize=normalize, asynchronous=out.is_cuda)

    def apply_T(self, out: torch.Tensor, y_block: torch.Tensor):
        """out = A_r^T y_r (NOT divided: the division follows the cross-rank sum)"""
        from .regridder import regrid_
        regrid_(out, self.RT, y_block, normalize=False, asynchronous=out.is_cuda)


class ShardedRegridder:
    """``Regridder(dst, src)`` sharded over the ranks of ``group`` by destination cells.

    ``local_factory(rows_grid, cols_grid) -> op`` builds the rank-local row-block operator (see
    :class:`_LocalB200` for the protocol); the default is the CUDA engine.  (The CPU tests inject a
    numpy/scipy factory to exercise the sharding and the collectives under gloo.)"""

    def __init__(self, dst: Grid, src: Grid, group=None, local_factory: Optional[Callable] = None,
                 device: Optional[torch.device] = None, normalize: bool = False):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_dst, self.n_src = dst.ncells, src.ncells
        self.device = device if device is not None else torch.device("cpu")
        factory = local_factory or _LocalB200
        self.dst_bounds = block_bounds(self.n_dst, self.world)
        lo, hi = self.dst_bounds[self.rank]
        self.local = factory(dst.slice(lo, hi), src)
        if normalize:
            # normalize!(R) (regridder.jl:54-62): A, dst_areas, src_areas ./= maximum(A); the maximum of a
            # row-sharded A is the max over the blocks' maxima -- one scalar all-reduce.
            m = torch.tensor([self.local.maximum()], dtype=torch.float64, device=self.device)
            if self.world > 1:
                dist.all_reduce(m, op=dist.ReduceOp.MAX, group=self.group)
            if float(m.item()) > 0.0:
                self.local.scale(float(m.item()))
        self._dst_areas = None
        self._src_areas = None
        self._nnz = None

    @property
    def shape(self):
        return (self.n_dst, self.n_src)

    # R.dst_areas / R.src_areas / nnz are collective on first access (all ranks must ask)
    @property
    def dst_areas(self) -> torch.Tensor:
        if self._dst_areas is None:
            self._dst_areas = self._all_gather_blocks(self.local.areas(self.device).to(self.device), self.dst_bounds)
        return self._dst_areas

    @property
    def src_areas(self) -> torch.Tensor:
        if self._src_areas is None:
            self._src_areas = self.local.src_areas(self.device).to(self.device)
        return self._src_areas

    @property
    def nnz(self) -> int:
        if self._nnz is None:
            t = torch.tensor([self.local.nnz], dtype=torch.int64, device=self.device)
            if self.world > 1:
                dist.all_reduce(t, group=self.group)
            self._nnz = int(t.item())
        return self._nnz

    # -- collectives -------------------------------------------------------------------------
    def _all_gather_blocks(self, shard: torch.Tensor, bounds) -> torch.Tensor:
        """Concatenate per-rank blocks (sizes differ by at most one) into the full vector(s).
        ``shard`` has the block rows in its first dimension."""
        n = bounds[-1][1]
        if self.world == 1:
            return shard.clone()
        width = max(hi - lo for lo, hi in bounds)
        if all(hi - lo == width for lo, hi in bounds):
            full = torch.empty((self.world * width,) + tuple(shard.shape[1:]), dtype=shard.dtype, device=shard.device)
            dist.all_gather_into_tensor(full, shard.contiguous(), group=self.group)
            return full
        pad = torch.zeros((width,) + tuple(shard.shape[1:]), dtype=shard.dtype, device=shard.device)
        pad[: shard.shape[0]] = shard
        full = torch.empty((self.world * width,) + tuple(shard.shape[1:]), dtype=shard.dtype, device=shard.device)
        dist.all_gather_into_tensor(full, pad, group=self.group)
        return torch.cat([full[k * width: k * width + (hi - lo)] for k, (lo, hi) in enumerate(bounds)])[:n]

    def _broadcast(self, x: Optional[torch.Tensor], n: int, trailing=(), root: int = 0) -> torch.Tensor:
        if self.world == 1:
            return x
        if x is None or self.rank != root:
            x = torch.empty((n,) + tuple(trailing), dtype=torch.float64, device=self.device)
        dist.broadcast(x, src=dist.get_global_rank(self.group, root) if self.group is not None else root,
                       group=self.group)
        return x

    # -- regrid! -------------------------------------------------------------------------------
    def regrid(self, field: Optional[torch.Tensor], transpose: bool = False, normalize: bool = True,
               broadcast: bool = True, gather: bool = True, trailing=()) -> torch.Tensor:
        """``regrid!`` (forward) or ``regrid!`` with ``transpose(R)``.

        forward: ``field`` (n_src, or (n_src, K) level-fastest) lives on rank 0 and is broadcast
        (``broadcast=False``: every rank already holds it); returns the all-gathered destination field
        (``gather=False``: this rank's block).
        transpose: ``field`` is the destination field -- full length on every rank, or this rank's
        block -- and the result is the full source-grid field on every rank (all-reduce)."""
        lo, hi = self.dst_bounds[self.rank]
        if not transpose:
            x = self._broadcast(field, self.n_src, trailing) if broadcast else field
            out = torch.zeros((hi - lo,) + tuple(x.shape[1:]), dtype=torch.float64, device=x.device)
            if hi > lo:
                self.local.apply(out, x, normalize)
            return self._all_gather_blocks(out, self.dst_bounds) if gather else out
        y = field
        y_block = y if y.shape[0] == hi - lo and self.world > 1 else y[lo:hi]
        part = torch.zeros((self.n_src,) + tuple(y.shape[1:]), dtype=torch.float64, device=y.device)
        if hi > lo:
            self.local.apply_T(part, y_block.contiguous())
        if self.world > 1:
            dist.all_reduce(part, group=self.group)
        if normalize:
            a = self.src_areas
            part /= a if part.dim() == 1 else a[:, None]
        return part
