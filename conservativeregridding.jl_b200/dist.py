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


def balanced_bounds(weights, world: int) -> List[Tuple[int, int]]:
    """Contiguous blocks of ``range(len(weights))`` with near-equal total weight (SURVEY.md section 8e:
    destination blocks balanced by candidate count).  ``weights``: non-negative numpy array or tensor."""
    w = torch.as_tensor(weights, dtype=torch.float64).flatten()
    n = int(w.numel())
    if world == 1 or n == 0:
        return block_bounds(n, world)
    c = torch.cumsum(w, 0)
    total = float(c[-1])
    if not total > 0.0:
        return block_bounds(n, world)
    targets = torch.arange(1, world, dtype=torch.float64, device=c.device) * (total / world)
    i = torch.searchsorted(c, targets).clamp_(max=n - 1)          # first cell whose cumulative weight reaches the target
    below = torch.where(i > 0, c[(i - 1).clamp_(min=0)], torch.zeros_like(targets))
    cuts = torch.where(c[i] - targets <= targets - below, i + 1, i).cpu().tolist()   # cut on the closer side
    edges = [0] + [min(max(int(x), 0), n) for x in cuts] + [n]
    for k in range(1, len(edges)):                    # monotone
        edges[k] = max(edges[k], edges[k - 1])
    return [(edges[k], edges[k + 1]) for k in range(world)]


def candidate_weights(dst: Grid, src: Grid):
    """Estimated broad-phase candidates per destination cell, (sqrt(a_dst) + sqrt(a_src))^2 / a_src with
    flat quad areas (fixed-stride 4-vertex cells only; None otherwise).  On cfg5 an equatorial 0.25 deg
    cell has ~10 candidates, a polar one ~3.5 -- equal-count latitude bands are 1.7x out of balance."""
    def flat_areas(g, stride=1):
        v = g.verts[::stride] if stride > 1 and g.offsets is None else g.verts
        if g.offsets is not None or not hasattr(v, "shape") or len(v.shape) != 3 or v.shape[1] != 4:
            return None
        v = torch.as_tensor(v)
        d1, d2 = v[:, 2] - v[:, 0], v[:, 3] - v[:, 1]
        if v.shape[2] == 3:
            return 0.5 * torch.linalg.cross(d1, d2).norm(dim=1)
        return 0.5 * (d1[:, 0] * d2[:, 1] - d1[:, 1] * d2[:, 0]).abs()
    if not isinstance(dst, Grid):
        return None
    ad = flat_areas(dst)
    if ad is None:
        return None
    if isinstance(src, Grid):
        a_s = flat_areas(src, max(1, src.ncells // 16384))       # the mean of a sample is enough
        if a_s is None or a_s.numel() == 0:
            return None
        a_s = float(a_s.mean())
    else:                                             # described grid: mean cell area of a global grid
        n_src = src.ncells
        a_s = 4.0 * np.pi * float(getattr(src, "radius", 1.0)) ** 2 / max(n_src, 1)
    if not a_s > 0.0:
        return None
    return (ad.sqrt() + a_s ** 0.5) ** 2 / a_s


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
    numpy/scipy factory to exercise the sharding and the collectives under gloo.)
    ``balance``: False = equal cell counts per block; True = blocks of equal estimated candidate count
    (:func:`candidate_weights`); an array = per-destination-cell weights.  Collective when not False.
    ``bounds``: explicit blocks (identical on every rank), e.g. ``dst_bounds`` of an earlier regridder."""

    def __init__(self, dst: Grid, src: Grid, group=None, local_factory: Optional[Callable] = None,
                 device: Optional[torch.device] = None, normalize: bool = False, balance=False,
                 bounds: Optional[List[Tuple[int, int]]] = None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_dst, self.n_src = dst.ncells, src.ncells
        self.device = device if device is not None else torch.device("cpu")
        factory = local_factory or _LocalB200
        self.dst_bounds = block_bounds(self.n_dst, self.world)
        if bounds is not None:                            # e.g. the dst_bounds of an earlier balanced regridder
            assert len(bounds) == self.world and bounds[0][0] == 0 and bounds[-1][1] == self.n_dst
            self.dst_bounds = [(int(lo), int(hi)) for lo, hi in bounds]
        elif balance and self.world > 1:
            # blocks of near-equal estimated work; rank 0 decides, everybody follows (bit-identical bounds)
            edges = torch.zeros(self.world + 1, dtype=torch.int64, device=self.device)
            if self.rank == 0:
                w = balance if not isinstance(balance, bool) else candidate_weights(dst, src)
                b = balanced_bounds(w, self.world) if w is not None else self.dst_bounds
                edges = torch.tensor([b[0][0]] + [hi for _, hi in b], dtype=torch.int64, device=self.device)
            dist.broadcast(edges, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                           group=self.group)
            e = edges.cpu().tolist()
            self.dst_bounds = [(int(e[k]), int(e[k + 1])) for k in range(self.world)]
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
