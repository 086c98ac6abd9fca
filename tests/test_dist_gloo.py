"""CPU test of the N>1 host path: destination sharding + collectives under gloo (world_size 2 and
3), with the rank-local operator supplied by the oracle instead of the CUDA engine."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _OracleLocal:
    """Row-block operator computed by the CPU oracle (test stand-in for the CUDA engine)."""

    def __init__(self, rows_grid, cols_grid):
        from oracle import oracle
        rows_grid, cols_grid = _explicit(rows_grid), _explicit(cols_grid)
        self.O = oracle.build_regridder(rows_grid, cols_grid)
        self.A = self.O.tocsc().tocsr()
        self.nnz = self.O.nnz

    def areas(self, device=None):
        return torch.from_numpy(self.O.dst_areas.copy())

    def src_areas(self, device=None):
        return torch.from_numpy(self.O.src_areas.copy())

    def maximum(self):
        return float(self.A.data.max()) if self.A.nnz else 0.0

    def scale(self, divisor):
        self.A = self.A / divisor
        self.O.dst_areas = self.O.dst_areas / divisor
        self.O.src_areas = self.O.src_areas / divisor

    def apply(self, out, x, normalize=True):
        y = self.A @ x.numpy()
        if normalize:
            y = y / (self.O.dst_areas if y.ndim == 1 else self.O.dst_areas[:, None])
        out.copy_(torch.from_numpy(np.asarray(y)))

    def apply_T(self, out, y_block, normalize=False):
        x = self.A.T @ y_block.numpy()
        if normalize:
            x = x / (self.O.src_areas if x.ndim == 1 else self.O.src_areas[:, None])
        out.copy_(torch.from_numpy(np.asarray(x)))


def _explicit(g):
    return g.materialize() if hasattr(g, "materialize") else g


def _oracle_areas(g):
    from oracle import oracle
    return oracle.cell_areas(_explicit(g))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from crg_b200 import grids
        from crg_b200.dist import ShardedRegridder, block_bounds
        from oracle import oracle
        dst, src = grids.lonlat_grid(25, 13), grids.healpix_grid(4, "ring")      # 325 cells: uneven blocks
        S = ShardedRegridder(dst, src, local_factory=_OracleLocal, areas_factory=_oracle_areas)
        full = oracle.build_regridder(dst, src)
        assert S.nnz == full.nnz and S.shape == (dst.ncells, src.ncells)
        assert np.allclose(S.dst_areas.numpy(), full.dst_areas, rtol=1e-15)
        assert np.allclose(S.src_areas.numpy(), full.src_areas, rtol=1e-15)
        x = torch.from_numpy(np.random.default_rng(0).random(src.ncells)) if rank == 0 else None
        y = S.regrid(x)                                   # broadcast from rank 0, all-gather
        x0 = np.random.default_rng(0).random(src.ncells)
        assert np.allclose(y.numpy(), full.regrid(x0), rtol=1e-13)
        xb = S.regrid(y, transpose=True)
        assert np.allclose(xb.numpy(), full.regrid(full.regrid(x0), transpose=True), rtol=1e-12)
        # batched (level-fastest) fields
        X = torch.from_numpy(np.stack([x0, 2 * x0, x0 ** 2], axis=1)) if rank == 0 else None
        Y = S.regrid(X, trailing=(3,))
        assert Y.shape == (dst.ncells, 3) and np.allclose(Y[:, 1].numpy(), 2 * full.regrid(x0), rtol=1e-13)
        # transpose from the local block only, and un-normalised (A^T y)
        lo, hi = block_bounds(dst.ncells, world)[rank]
        xb2 = S.regrid(y[lo:hi].clone(), transpose=True) if world > 1 else xb
        assert np.allclose(xb2.numpy(), xb.numpy(), rtol=1e-14)
        xr = S.regrid(y, transpose=True, normalize=False)
        assert np.allclose(xr.numpy(), full.tocsc().T @ y.numpy(), rtol=1e-12)
        # local block only
        lo, hi = block_bounds(dst.ncells, world)[rank]
        yl = S.regrid(torch.from_numpy(x0), broadcast=False, gather=False)
        assert yl.shape[0] == hi - lo and np.allclose(yl.numpy(), full.regrid(x0)[lo:hi], rtol=1e-13)
        # blocks balanced by estimated candidate count (rank 0 decides, bounds broadcast): same results
        Sb = ShardedRegridder(dst, src, local_factory=_OracleLocal, areas_factory=_oracle_areas, balance=True)
        assert Sb.dst_bounds[0][0] == 0 and Sb.dst_bounds[-1][1] == dst.ncells
        assert all(a[1] == b[0] for a, b in zip(Sb.dst_bounds[:-1], Sb.dst_bounds[1:]))
        if world == 3:                                         # (two blocks of a symmetric grid are balanced already)
            assert Sb.dst_bounds != S.dst_bounds             # polar rows are cheaper than equatorial ones
        assert np.allclose(Sb.dst_areas.numpy(), full.dst_areas, rtol=1e-15)
        yb = Sb.regrid(torch.from_numpy(x0), broadcast=False)
        assert np.allclose(yb.numpy(), full.regrid(x0), rtol=1e-13)
        xbb = Sb.regrid(yb, transpose=True)
        assert np.allclose(xbb.numpy(), xb.numpy(), rtol=1e-12)
        # normalize!(R): every block scaled by the global maximum(A) (one scalar all-reduce)
        Sn = ShardedRegridder(dst, src, local_factory=_OracleLocal, areas_factory=_oracle_areas, normalize=True)
        m = full.tocsc().max()
        assert np.allclose(Sn.dst_areas.numpy(), full.dst_areas / m, rtol=1e-15)
        assert np.allclose(Sn.src_areas.numpy(), full.src_areas / m, rtol=1e-15)
        assert abs(max(Sn.local.maximum(), 0.0) - (full.tocsc()[lo:hi].max() / m)) < 1e-15
        yn = Sn.regrid(torch.from_numpy(x0), broadcast=False)
        assert np.allclose(yn.numpy(), full.regrid(x0), rtol=1e-13)          # the normalisation cancels in regrid!
        # halo-sliced sources: described grids (ring tables) and explicit cells (per-cell latitude ranges); every rank
        # builds against a strict subset of the source, the transpose is all-gather + overlap-add (no reduction)
        dspec, sspec = grids.lonlat_spec(60, 30), grids.healpix_spec(16, "ring")
        fullh = oracle.build_regridder(dspec.materialize(), sspec.materialize())
        xh = np.random.default_rng(1).random(sspec.ncells)
        for d_, s_ in ((dspec, sspec), (dspec.materialize(), sspec.materialize()), (dspec, sspec.materialize())):
            Sh = ShardedRegridder(d_, s_, local_factory=_OracleLocal, areas_factory=_oracle_areas)
            a, b = Sh.src_range
            assert 0 <= a < b <= sspec.ncells and (b - a) < sspec.ncells, (a, b)
            assert Sh.nnz == fullh.nnz
            assert np.allclose(Sh.src_areas.numpy(), fullh.src_areas, rtol=1e-15)
            yh = Sh.regrid(torch.from_numpy(xh), broadcast=False)
            assert np.allclose(yh.numpy(), fullh.regrid(xh), rtol=1e-13)
            xbh = Sh.regrid(yh, transpose=True)
            assert np.allclose(xbh.numpy(), fullh.regrid(fullh.regrid(xh), transpose=True), rtol=1e-12)
            part = Sh.regrid(yh, transpose=True, gather=False)
            assert part.shape[0] == b - a
        # the halo must contain every source cell that meets the block (else entries would be missing): nnz above;
        # without halos the same results
        Sn0 = ShardedRegridder(dspec, sspec, local_factory=_OracleLocal, areas_factory=_oracle_areas, halo=False)
        assert Sn0.src_range == (0, sspec.ncells)
        assert np.allclose(Sn0.regrid(torch.from_numpy(xh), broadcast=False).numpy(), fullh.regrid(xh), rtol=1e-13)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_regridder_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_block_bounds():
    from crg_b200.dist import block_bounds
    assert block_bounds(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert block_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    b = block_bounds(3145728, 8)
    assert b[0] == (0, 393216) and b[-1][1] == 3145728
