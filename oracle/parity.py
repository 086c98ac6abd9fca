"""Comparison of two regridder matrices in the terms of north_star (TEST INFRASTRUCTURE, like the rest of oracle/:
used by tests/ and by bench.py's cpu_baseline leg to report `parity`)."""
import numpy as np


def pattern_threshold(dst_areas, src_areas):
    """Entries below this are round-off slivers of edge-coincident cells:
    the reference keeps `area > 0`, whose outcome on such pairs is rounding noise."""
    pos = np.concatenate([dst_areas[dst_areas > 0], src_areas[src_areas > 0]])
    return 1e-9 * float(pos.min())



def parity_report(A, B, dst_areas, src_areas, rtol=1e-10, symdiff_out=None):
    """Per-entry comparison of two sparse matrices (A = device, B = oracle) in the terms of north_star:
    how many entries each side has below the sliver threshold tau, the symmetric difference of the
    patterns and its largest value, the pattern difference ABOVE tau (must be 0), the largest relative and
    absolute error over the common entries, and how many exceed `rtol` relative + the absolute floor.
    Works on key arrays (no sparse subtraction): fine for the 9.3 M entries of BASELINE config 5."""
    A = A.tocsc(); B = B.tocsc()
    A.sort_indices(); B.sort_indices()
    n_dst = A.shape[0]
    ka = np.repeat(np.arange(A.shape[1], dtype=np.int64), np.diff(A.indptr)) * n_dst + A.indices
    kb = np.repeat(np.arange(B.shape[1], dtype=np.int64), np.diff(B.indptr)) * n_dst + B.indices
    tau = pattern_threshold(dst_areas, src_areas)
    both, ia, ib = np.intersect1d(ka, kb, assume_unique=True, return_indices=True)
    only_a = np.ones(ka.size, dtype=bool); only_a[ia] = False
    only_b = np.ones(kb.size, dtype=bool); only_b[ib] = False
    va, vb = A.data[ia], B.data[ib]
    d = np.abs(va - vb)
    floor = 1e-12 * float(B.data.max()) if B.nnz else 0.0
    big = np.maximum(va, vb) > tau
    rel = np.where(big, d / np.maximum(vb, 1e-300), 0.0)
    bad = big & (d > rtol * vb + floor)
    sym_vals = np.concatenate([A.data[only_a], B.data[only_b]])
    if symdiff_out is not None:          # keys col * n_dst + row of the entries only one side keeps
        symdiff_out["only_device"], symdiff_out["only_oracle"] = ka[only_a], kb[only_b]
        symdiff_out["only_device_val"], symdiff_out["only_oracle_val"] = A.data[only_a], B.data[only_b]
    return {
        "tau": tau, "floor": floor, "nnz_device": int(A.nnz), "nnz_oracle": int(B.nnz),
        "n_under_tau_device": int((A.data <= tau).sum()), "n_under_tau_oracle": int((B.data <= tau).sum()),
        "n_only_device": int(only_a.sum()), "n_only_oracle": int(only_b.sum()),
        "symdiff_max_value": float(sym_vals.max()) if sym_vals.size else 0.0,
        "n_pattern_diff_above_tau": int((sym_vals > tau).sum()),
        "max_rel": float(rel.max()) if rel.size else 0.0, "max_abs": float(d.max()) if d.size else 0.0,
        "n_entries_beyond_tolerance": int(bad.sum()),
        # relative error of the entries above the absolute floor only (slivers are dominated by the 1e-16 position
        # round-off of their vertices, whatever the implementation)
        "max_rel_above_floor": float(np.where(d > floor, rel, 0.0).max()) if rel.size else 0.0,
    }
