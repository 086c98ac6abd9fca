"""Generates tests/golden/highprec_pairs.npz: >= 10^4 (source cell, destination cell) pairs sampled from the
candidate lists of BASELINE configs 5, 2 and 1 THEMSELVES, with their intersection area computed by an
independent algorithm in exact / 50-digit arithmetic (oracle/highprec.py::independent_intersection_area:
exact integer in/out predicates, vertex enumeration, angular sort, Girard's excess -- no Sutherland-Hodgman,
no Float64).  The fixture stores the pairs' Float64 vertices, so the checks need neither grids.py nor mpmath:

  * tests/test_oracle.py           pins the ORACLE's Float64 clip on these pairs          (CPU)
  * tests/test_gpu_highprec.py     pins the CUDA clip (crg_clip_pairs) on the same pairs  (GPU)

Categories: random candidates, polar cells, the smallest positive overlaps (slivers), candidates without
overlap (false positives of the broad phase, incl. touching cells), pairs with a vertex exactly on a
coordinate plane (HEALPix corners on lon-lat lines), exactly nested / edge-coincident cells (config 1), and --
when gpurun_out/symdiff_*.npz from a GPU run of tests/test_gpu_parity_full.py is present -- the pairs that only
one of (CUDA, oracle) keeps.

Run (about a minute on 8 cores):  python tests/golden/make_highprec_pairs.py
"""
import glob
import os
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from crg_b200 import grids  # noqa: E402
from oracle import oracle  # noqa: E402

CONFIGS = {
    "cfg5": (lambda: grids.lonlat_grid(1440, 720), lambda: grids.healpix_grid(512, "ring"),
             dict(random=4500, polar=1200, sliver=1200, zero=800, aligned=800)),
    "cfg2": (lambda: grids.lonlat_grid(720, 360), lambda: grids.healpix_grid(256, "ring"),
             dict(random=2000, polar=500, sliver=500, zero=300, aligned=0)),
    "cfg1": (lambda: grids.lonlat_grid(180, 90), lambda: grids.lonlat_grid(360, 180),
             dict(random=600, polar=100, sliver=0, zero=100, aligned=0)),
}
CATS = ["random", "polar", "sliver", "zero", "aligned", "symdiff"]


def _area(args):
    from oracle import highprec
    return highprec.independent_intersection_area(args[0], args[1])


def sample(name, dst, src, counts, rng):
    nthreads = oracle.use_all_cores()
    ps, pd = oracle.dual_query(oracle.treeify(src), oracle.treeify(dst), nthreads)
    order = np.lexsort((ps, pd))
    ps, pd = ps[order], pd[order]
    # oracle area of every candidate (0 where dropped)
    i1, i2, a = oracle.compute_intersection_areas(dst, src, ps, pd, nthreads)
    key_all = pd * src.ncells + ps
    area = np.zeros(len(ps))
    area[np.searchsorted(key_all, i2 * src.ncells + i1)] = a
    chosen = {}

    def take(cat, idx, n):
        idx = np.setdiff1d(idx, np.fromiter(chosen.keys(), dtype=np.int64, count=len(chosen)))
        if n and len(idx):
            for k in rng.choice(idx, size=min(n, len(idx)), replace=False):
                chosen[int(k)] = CATS.index(cat)

    nd = dst.ncells
    nx = dst.meta["shape"][0]
    polar_dst = (pd < nx) | (pd >= nd - nx)
    polar_src = (ps < 8) | (ps >= src.ncells - 8)
    take("polar", np.nonzero(polar_dst | polar_src)[0], counts["polar"])
    pos = np.nonzero(area > 0)[0]
    # slivers: the smallest positive overlaps -- half of them round-off residues of touching cells (the entries
    # whose presence in the pattern is noise), half the smallest overlaps above the pattern threshold tau
    amin = min(oracle.cell_areas(dst).min(), oracle.cell_areas(src)[oracle.cell_areas(src) > 0].min())
    take("sliver", pos[np.argsort(area[pos])[: 2 * counts["sliver"]]], counts["sliver"] // 2)
    real = pos[area[pos] > 1e-9 * amin]
    take("sliver", real[np.argsort(area[real])[: 2 * counts["sliver"]]], counts["sliver"] - counts["sliver"] // 2)
    take("zero", np.nonzero(area == 0)[0], counts["zero"])
    if counts["aligned"]:
        onplane = (src.verts == 0.0).any(axis=(1, 2))
        take("aligned", np.nonzero(onplane[ps])[0], counts["aligned"])
    for fn in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"symdiff_{name}*.npz")) +
                     glob.glob(os.path.join(HERE, f"symdiff_{name}*.npz"))):
        z = np.load(fn)
        k = np.concatenate([z["only_device"], z["only_oracle"]]).astype(np.int64)     # keys col * n_dst + row
        s_, d_ = k // nd, k % nd
        pos_ = np.searchsorted(key_all, d_ * src.ncells + s_)
        ok = (pos_ < len(key_all)) & (key_all[np.minimum(pos_, len(key_all) - 1)] == d_ * src.ncells + s_)
        for q in pos_[ok][:1500]:
            chosen.setdefault(int(q), CATS.index("symdiff"))
        print(f"  {fn}: {int(ok.sum())} of {len(k)} symmetric-difference pairs are oracle candidates")
    take("random", np.arange(len(ps)), counts["random"])
    idx = np.array(sorted(chosen), dtype=np.int64)
    cat = np.array([chosen[int(k)] for k in idx], dtype=np.int8)
    return src.verts[ps[idx]], dst.verts[pd[idx]], ps[idx], pd[idx], cat, area[idx], amin


def main():
    rng = np.random.default_rng(20260102)
    out = {k: [] for k in ("src_verts", "dst_verts", "src_idx", "dst_idx", "cat", "cfg", "oracle_area", "min_cell_area")}
    for c, (name, (fd, fs, counts)) in enumerate(CONFIGS.items()):
        dst, src = fd(), fs()
        sv, dv, si, di, cat, oa, amin = sample(name, dst, src, counts, rng)
        print(name, len(si), "pairs", {CATS[k]: int((cat == k).sum()) for k in range(len(CATS))})
        for k, v in zip(out, (sv, dv, si, di, cat, np.full(len(si), c, dtype=np.int8), oa, np.full(len(si), amin))):
            out[k].append(v)
    out = {k: np.concatenate(v) for k, v in out.items()}
    with Pool() as pool:
        exact = np.array(pool.map(_area, list(zip(out["src_verts"], out["dst_verts"])), chunksize=64))
    out["area"] = exact
    del out["oracle_area"]          # the fixture holds the independent answer only
    fn = os.path.join(HERE, "highprec_pairs.npz")
    np.savez_compressed(fn, categories=np.array(CATS), configs=np.array(list(CONFIGS)), **out)
    print(fn, len(exact), "pairs,", os.path.getsize(fn), "bytes; zero-area pairs:", int((exact == 0).sum()))


if __name__ == "__main__":
    main()
