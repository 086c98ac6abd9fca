"""Destination-sharded multi-GPU regridder (one process per GPU, ``torch.distributed``).

SURVEY.md section 8(e): destination cells are split into contiguous field-index blocks, one per
rank; each rank builds the row block ``A_r`` of ``A`` for its destination cells against the HALO of
the source grid that can reach its block -- ONE local build, no exchange during the build.

* halo: the source cells whose latitude range meets the (inflated) latitude range of the block form a
  contiguous index range for ring-major source grids (HEALPix ring, lon-lat, RingGrids full / octahedral);
  described grids (:class:`GridSpec`) generate exactly that range on the device, so the per-rank cost of
  the source side (cell bounds, binning, the rows of ``A_r^T``) shrinks with the number of ranks.  Other
  orders (nested, cubed sphere) fall back to the replicated source + the build's own culling box.
* forward ``regrid!``: the source field is broadcast (NCCL), every rank computes its block
  ``y_r = (A_r x[halo]) ./ a_dst_r`` and the blocks are all-gathered;
* ``transpose(R)``: ``A^T y = sum_r A_r^T y_r`` -- every rank applies the transpose of its own block
  (the CSR(A_r^T) the local assembly produces anyway, division by the source areas fused: the areas are per
  source cell, so dividing before the sum is the same) to its slice of the destination field; the partial
  vectors only cover the rank's halo, they are ALL-GATHERED (no reduction collective) and overlap-added.
* ``R.dst_areas``: all-gather of the per-block areas; ``R.src_areas``: every rank computes an equal share
  (``crg_grid_areas``), all-gathered.
* ``normalize=True``: every block is scaled by the all-reduced (max) ``maximum(A_r)`` (one scalar).

The reference has no distributed path at all; the single-process semantics
this reproduces are ``Regridder`` / ``regrid!`` / ``transpose`` (src/regridder/regridder.jl:125-163,
49-50; src/regridder/regrid.jl:63-118).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .grids import Grid, GridSpec, ring_table


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


def _z_range(block):
    """(zmin, zmax) of the vertices of a destination block (Grid with numpy / torch vertices, or GridSpec)."""
    if isinstance(block, GridSpec):
        t = ring_table(block)
        if t is None:
            return None
        start, zlo, zhi = t
        lo, hi = (block.cell_lo, block.cell_hi) if (block.cell_lo or block.cell_hi) else (0, block.ncells_full)
        if hi <= lo:
            return None
        r0 = int(np.searchsorted(start, lo, side="right")) - 1
        r1 = int(np.searchsorted(start, hi - 1, side="right")) - 1
        return float(min(zlo[r0:r1 + 1].min(), zhi[r0:r1 + 1].min())), float(max(zlo[r0:r1 + 1].max(), zhi[r0:r1 + 1].max()))
    if block.manifold != 1 or block.ncells == 0:
        return None
    z = block.verts[..., 2]
    return float(z.min()), float(z.max())


_HALO_CACHE = {}


def _spec_key(g: GridSpec):
    return (g.kind, g.n1, g.n2, tuple(g.p), g.flags, g.cell_lo, g.cell_hi,
            None if g.lat_deg is None else hash(np.asarray(g.lat_deg).tobytes()))


def halo_range(block, src) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of source cells (indices of ``src`` as given) that can intersect the destination
    ``block``: every source cell whose vertex latitude range, inflated by the bulge of great-circle edges over a
    cell, meets the block's.  Tight for ring-major source orders; (0, n_src) when nothing is known.
    (For two described grids the answer is a property of the descriptors and is remembered.)"""
    if isinstance(block, GridSpec) and isinstance(src, GridSpec):
        key = (_spec_key(block), _spec_key(src))
        if key not in _HALO_CACHE:
            if len(_HALO_CACHE) > 4096:
                _HALO_CACHE.clear()
            _HALO_CACHE[key] = _halo_range(block, src)
        return _HALO_CACHE[key]
    return _halo_range(block, src)


def _halo_range(block, src) -> Tuple[int, int]:
    n_src = src.ncells
    zr = _z_range(block)
    if zr is None or n_src == 0:
        return 0, n_src
    # an edge of length d bulges by at most d^2 / 8 in z beyond its end points; d < 4 sqrt(4 pi / n) for the grids here
    nd_full = block.ncells_full if isinstance(block, GridSpec) else None
    d_dst = 4.0 * np.sqrt(4.0 * np.pi / max(nd_full or 0, 1)) if nd_full else None
    if d_dst is None:                                     # explicit block: its own largest cell diameter
        v = block.verts
        if block.offsets is not None:
            return 0, n_src
        diff = v[:, :, None, :] - v[:, None, :, :] if v.shape[0] <= 4096 else v[::max(1, v.shape[0] // 4096)][:, :, None, :] - v[::max(1, v.shape[0] // 4096)][:, None, :, :]
        d_dst = 2.0 * float((diff * diff).sum(-1).max()) ** 0.5
    ns_full = src.ncells_full if isinstance(src, GridSpec) else src.ncells
    d_src = 4.0 * np.sqrt(4.0 * np.pi / max(ns_full, 1))
    margin = d_dst * d_dst + d_src * d_src + 1e-9
    zmin, zmax = zr[0] - margin, zr[1] + margin
    if isinstance(src, GridSpec):
        t = ring_table(src)
        if t is None:
            return 0, n_src
        start, zlo, zhi = t
        hit = np.nonzero((zhi >= zmin) & (zlo <= zmax))[0]
        if hit.size == 0:
            return 0, 0
        lo, hi = int(start[hit[0]]), int(start[hit[-1] + 1])
        base = src.cell_lo if (src.cell_lo or src.cell_hi) else 0
        top = src.cell_hi if (src.cell_lo or src.cell_hi) else src.ncells_full
        lo, hi = max(lo, base), min(hi, top)
        return (lo - base, max(hi, lo) - base)
    if src.manifold != 1 or src.offsets is not None:
        return 0, n_src
    z = src.verts[..., 2]
    hit = ((z.max(-1) >= zmin) & (z.min(-1) <= zmax)).nonzero()
    hit = hit[0] if isinstance(hit, tuple) else hit.flatten()
    if len(hit) == 0:
        return 0, 0
    return int(hit[0]), int(hit[-1]) + 1


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

    def apply_T(self, out: torch.Tensor, y_block: torch.Tensor, normalize: bool = False):
        """out = A_r^T y_r, divided by the (halo) source areas when ``normalize`` (per source cell, so dividing
        each rank's partial vector before the cross-rank sum equals dividing the sum)"""
        from .regridder import regrid_
        regrid_(out, self.RT, y_block, normalize=normalize, asynchronous=out.is_cuda)


class ShardedRegridder:
    """``Regridder(dst, src)`` sharded over the ranks of ``group`` by destination cells.

    ``local_factory(rows_grid, cols_grid) -> op`` builds the rank-local row-block operator (see
    :class:`_LocalB200` for the protocol); the default is the CUDA engine.  (The CPU tests inject a
    numpy/scipy factory to exercise the sharding and the collectives under gloo.)
    ``balance``: False = equal cell counts per block; True = blocks of equal estimated candidate count
    (:func:`candidate_weights`); an array = per-destination-cell weights.  Collective when not False.
    ``bounds``: explicit blocks (identical on every rank), e.g. ``dst_bounds`` of an earlier regridder.
    ``halo``: build against the source halo of the block (:func:`halo_range`) instead of the replicated source.
    ``areas_factory(grid) -> tensor``: geometric cell areas of a grid slice (default: ``crg_grid_areas``)."""

    def __init__(self, dst: Grid, src: Grid, group=None, local_factory: Optional[Callable] = None,
                 device: Optional[torch.device] = None, normalize: bool = False, balance=False,
                 bounds: Optional[List[Tuple[int, int]]] = None, halo: bool = True,
                 areas_factory: Optional[Callable] = None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_dst, self.n_src = dst.ncells, src.ncells
        self.device = device if device is not None else torch.device("cpu")
        factory = local_factory or _LocalB200
        self._areas_factory = areas_factory
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
        block = dst.slice(lo, hi)
        self.src = src
        self.src_range = halo_range(block, src) if (halo and self.world > 1) else (0, self.n_src)
        self.local = factory(block, src.slice(*self.src_range) if self.src_range != (0, self.n_src) else src)
        # every rank's halo range: a function of the descriptors for described grids (computed locally, remembered),
        # collective on the first transpose otherwise
        self._ranges = None
        if not (halo and self.world > 1):
            self._ranges = [(0, self.n_src)] * self.world
        elif isinstance(dst, GridSpec) and isinstance(src, GridSpec):
            self._ranges = [halo_range(dst.slice(a, b), src) for a, b in self.dst_bounds]
        self._scale = None
        if normalize:
            # normalize!(R) (regridder.jl:54-62): A, dst_areas, src_areas ./= maximum(A); the maximum of a
            # row-sharded A is the max over the blocks' maxima -- one scalar all-reduce.
            m = torch.tensor([self.local.maximum()], dtype=torch.float64, device=self.device)
            if self.world > 1:
                dist.all_reduce(m, op=dist.ReduceOp.MAX, group=self.group)
            if float(m.item()) > 0.0:
                self.local.scale(float(m.item()))
                self._scale = float(m.item())
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
        """Geometric areas of ALL source cells (collective on first access): an equal share per rank, all-gathered."""
        if self._src_areas is None:
            if self.src_range == (0, self.n_src) and self.world == 1:
                self._src_areas = self.local.src_areas(self.device).to(self.device)
            else:
                shares = block_bounds(self.n_src, self.world)
                lo, hi = shares[self.rank]
                mine = self._cell_areas(self.src.slice(lo, hi))
                self._src_areas = self._all_gather_blocks(mine, shares)
                if self._scale is not None:
                    self._src_areas = self._src_areas / self._scale
        return self._src_areas

    def _cell_areas(self, grid) -> torch.Tensor:
        if self._areas_factory is not None:
            return torch.as_tensor(self._areas_factory(grid), dtype=torch.float64).to(self.device)
        from .regridder import areas
        out = torch.empty(grid.ncells, dtype=torch.float64, device=self.device)
        if out.is_cuda:
            from .regridder import torch_stream_ptr
            areas(grid, out=out, device=self.device.index, stream=torch_stream_ptr(self.device))
            return out
        return torch.from_numpy(areas(grid))

    def halo_ranges(self) -> List[Tuple[int, int]]:
        """Every rank's source halo range (collective on first call)."""
        if self._ranges is None:
            t = torch.tensor(list(self.src_range), dtype=torch.int64, device=self.device)
            if self.world > 1:
                allr = torch.empty(2 * self.world, dtype=torch.int64, device=self.device)
                dist.all_gather_into_tensor(allr, t, group=self.group)
                v = allr.cpu().tolist()
            else:
                v = t.cpu().tolist()
            self._ranges = [(int(v[2 * k]), int(v[2 * k + 1])) for k in range(self.world)]
        return self._ranges

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
        s_lo, s_hi = self.src_range
        if not transpose:
            x = self._broadcast(field, self.n_src, trailing) if broadcast else field
            if hi > lo and s_hi > s_lo:
                out = torch.empty((hi - lo,) + tuple(x.shape[1:]), dtype=torch.float64, device=x.device)
                self.local.apply(out, x[s_lo:s_hi], normalize)           # the halo rows of x: a contiguous view
            else:
                out = torch.zeros((hi - lo,) + tuple(x.shape[1:]), dtype=torch.float64, device=x.device)
            return self._all_gather_blocks(out, self.dst_bounds) if gather else out
        y = field
        y_block = y if y.shape[0] == hi - lo and self.world > 1 else y[lo:hi]
        # (the buffer is as long as the LONGEST halo, so that it can go into the all-gather as it is)
        width = max(b - a for a, b in self._ranges) if (self._ranges is not None and gather and self.world > 1) else s_hi - s_lo
        buf = torch.empty((width,) + tuple(y.shape[1:]), dtype=torch.float64, device=y.device)
        part = buf[: s_hi - s_lo]
        if hi > lo and s_hi > s_lo:
            self.local.apply_T(part, y_block.contiguous(), normalize)
        else:
            part.zero_()
        if self.world == 1:
            if (s_lo, s_hi) == (0, self.n_src):
                return part
            full = torch.zeros((self.n_src,) + tuple(y.shape[1:]), dtype=torch.float64, device=y.device)
            full[s_lo:s_hi] = part
            return full
        if not gather:
            return part                                                 # covers source cells self.src_range
        # all-gather of the halo partials + overlap-add: no reduction collective (the halos of neighbouring
        # blocks share a few rings, everything else is written once)
        ranges = self.halo_ranges()
        pad = buf
        if buf.shape[0] != max(b - a for a, b in ranges):          # (ranges were not known before the apply)
            width = max(b - a for a, b in ranges)
            pad = torch.empty((width,) + tuple(part.shape[1:]), dtype=torch.float64, device=part.device)
            pad[: part.shape[0]] = part
        allp = torch.empty((self.world * width,) + tuple(part.shape[1:]), dtype=torch.float64, device=part.device)
        dist.all_gather_into_tensor(allp, pad, group=self.group)
        # ranges of ring-major halos are increasing with the rank: block k is COPIED where nothing was written yet and
        # ADDED where it overlaps its predecessors; source cells no block reaches stay zero
        full = torch.empty((self.n_src,) + tuple(y.shape[1:]), dtype=torch.float64, device=y.device)
        done = 0                                         # cells [0, done) are written
        order = sorted(range(self.world), key=lambda k: ranges[k])
        monotone = all(ranges[order[i]][1] <= ranges[order[i + 1]][1] for i in range(self.world - 1))
        if not monotone:
            full.zero_()
            for k, (a, b) in enumerate(ranges):
                if b > a:
                    full[a:b] += allp[k * width: k * width + (b - a)]
            return full
        for k in order:
            a, b = ranges[k]
            if b <= a:
                continue
            blk = allp[k * width: k * width + (b - a)]
            if a > done:
                full[done:a].zero_()
            ov = max(min(done, b) - a, 0)                # [a, a + ov) already written: add
            if ov > 0:
                full[a:a + ov] += blk[:ov]
            if b > a + ov:
                full[a + ov:b] = blk[ov:]
            done = max(done, b)
        if done < self.n_src:
            full[done:].zero_()
        return full
