#!/usr/bin/env python
"""bench.py -- headline benchmark of the two hot paths on B200 (see DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg5|cfg4|cfg3|cfg2|cfg1]

One *step* = one pass of the hot path over the workload: build the Regridder (broad phase ->
clip/area -> sort/assemble CSR + CSC -> areas), then regrid! forward and regrid! with
transpose(R).  Default workload = BASELINE.json configs[4]: 0.25 deg lon-lat (1440 x 720,
destination) <-> HEALPix nside=512 ring (source); it fits one GPU.  For N > 1 the destination
cells are sharded over the ranks, every rank builds against the source HALO of its block ("strong"
scaling: total work fixed), fields are broadcast / all-gathered with NCCL (no reduction collective).

metric = overlapping cell pairs (= nnz of the regridder, identical for every implementation)
processed per second through the whole step.  `value`: grids and fields already resident in
HBM; `e2e`: host (pinned) buffers in, host buffers out, copies inside the timed region.
`--impl reference` times the CPU restatement of the reference algorithm (oracle/, OpenMP on all
host cores) on the same workload: Julia is not installed here, so this is kind="port".
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "regridder_build_plus_regrid_overlapping_cell_pairs_per_s"
UNIT = "cell-pairs/s"

# name: (description, dst spec factory, src spec factory, K = fields per regrid! call)
WORKLOADS = {
    "cfg5": ("0.25deg lon-lat 1440x720 (dst) <-> HEALPix nside=512 ring (src): build + regrid! fwd + transpose",
             lambda g: g.lonlat_spec(1440, 720), lambda g: g.healpix_spec(512, "ring"), 1),
    "cfg5x4": ("4x config 5 in every count: 0.125deg lon-lat 2880x1440 (dst) <-> HEALPix nside=1024 ring (src): build + regrid! fwd + transpose",
               lambda g: g.lonlat_spec(2880, 1440), lambda g: g.healpix_spec(1024, "ring"), 1),
    "cfg4": ("full Gaussian F160 640x320 (dst) <-> octahedral Gaussian O320 (src): build + regrid! fwd + transpose",
             lambda g: g.full_gaussian_spec(160), lambda g: g.octahedral_gaussian_spec(320), 1),
    "cfg3": ("1deg lon-lat 360x180 (dst) <- equiangular cubed sphere C180 (src), 100-level field (dims=1, cell-fastest): "
             "build + batched regrid! fwd + transpose",
             lambda g: g.lonlat_spec(360, 180), lambda g: g.cubed_sphere_spec(180), 100),
    "cfg2": ("0.5deg lon-lat 720x360 (dst) <-> HEALPix nside=256 ring (src): build + regrid! fwd + transpose",
             lambda g: g.lonlat_spec(720, 360), lambda g: g.healpix_spec(256, "ring"), 1),
    "cfg1": ("2deg lon-lat 180x90 (dst) <- 1deg lon-lat 360x180 (src): build + regrid! fwd + transpose",
             lambda g: g.lonlat_spec(180, 90), lambda g: g.lonlat_spec(360, 180), 1),
}

COUNTERS_JSON = os.path.join(ROOT, "profiles", "r02_kernel_counters.json")
CSRC = os.path.join(ROOT, "conservativeregridding.jl_b200", "csrc")


def source_hash(files):
    h = hashlib.sha256()
    for f in files:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def kernel_counters():
    """FP64 work per candidate pair of the clip kernel and DRAM traffic per launch, as MEASURED by ncu on the shipped
    kernels (profiles/r02_kernel_counters.json, written by scripts/ncu_summary.py --json from the .ncu-rep files).
    The entry is only used when the kernel sources still hash to what was profiled; otherwise it is flagged stale."""
    out = {"clip": None, "spmv_fwd": None, "spmv_T": None}
    try:
        d = json.load(open(COUNTERS_JSON))
    except Exception:
        return out, "missing " + os.path.relpath(COUNTERS_JSON, ROOT)
    stale = []
    for k in out:
        e = d.get(k)
        if not e:
            continue
        if e.get("src_hash") == source_hash(e.get("src_files", [])):
            out[k] = e
        else:
            stale.append(k)
    return out, ("stale: " + ",".join(stale)) if stale else "ok"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs.  One long-running
    `nvidia-smi -lms` process, started BEFORE the warm-up steps: its start-up (fork + NVML init, which
    takes driver locks) stalled the first timed build by several milliseconds when it was started at the
    beginning of the timed region.  Only the rows read inside the timed region are reported."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (t, r) in self.rows if t0 is None or (t0 <= t <= t1 + 0.05)]
        if not rows and self.rows:          # region shorter than one polling interval: the closest row
            rows = [min(self.rows, key=lambda tr: abs(tr[0] - (t1 if t1 is not None else tr[0])))[1]]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_field(n, K, rng):
    """K source fields of n cells, cell-fastest (the reference's dims=1 layout of an (n, K) Julia array)."""
    return rng.random(n) if K == 1 else rng.random((K, n))


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the restated reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------

def cpu_reference_step(oracle, dst, src, trees, x, nthreads):
    """One step on the CPU: dual-DFS candidates (threaded) -> per-pair clip + area (threaded) ->
    serial sparse() -> serial areas -> serial mul! forward and transposed, one SpMV per level like the
    reference's NDSliceLoop (the reference's own threading model).  Tree construction is
    excluded (the reference's trees are lazy, O(1)).  Grids without a reference tree (octahedral: RingGridsExt.jl:18-20
    errors) take k-d tree candidates instead of the dual DFS (a FlatNoTree would be O(N M))."""
    t0 = time.perf_counter()
    cands = oracle.dual_query(trees[1], trees[0], nthreads) if trees is not None else None
    R = oracle.build_regridder(dst, src, candidates=cands, nthreads=nthreads)
    t1 = time.perf_counter()
    y = R.regrid(x) if x.ndim == 1 else np.stack([R.regrid(xk) for xk in x])
    t2 = time.perf_counter()
    _ = R.regrid(y, transpose=True) if x.ndim == 1 else np.stack([R.regrid(yk, transpose=True) for yk in y])
    t3 = time.perf_counter()
    return R, y, (t1 - t0, t2 - t1, t3 - t2)


def oracle_trees(oracle, dst, src):
    if all(g.meta.get("kind") in ("lonlat", "full_ring", "healpix", "cubed_sphere") for g in (dst, src)):
        return (oracle.treeify(dst), oracle.treeify(src))
    return None


def base_config(desc, n_dst, n_src, K):
    """`config` is identical in both arms (the driver compares them)."""
    return {"workload": desc, "n_dst": n_dst, "n_src": n_src, "fields_per_regrid": K}


def run_reference(args, rank):
    if rank != 0:
        return
    from crg_b200 import grids
    from oracle import oracle
    oracle.build()
    desc, fd, fs, K = WORKLOADS[args.workload]
    dst, src = fd(grids).materialize(), fs(grids).materialize()
    nthreads = oracle.use_all_cores()          # torchrun exports OMP_NUM_THREADS=1
    trees = oracle_trees(oracle, dst, src)
    x = make_field(src.ncells, K, np.random.default_rng(20260101))
    times = []
    R = None
    for it in range(args.warmup + args.steps):
        R, _, t = cpu_reference_step(oracle, dst, src, trees, x, nthreads)
        if it >= args.warmup:
            times.append(t)
    tot = sum(sum(t) for t in times)
    ms = 1e3 * tot / len(times)
    value = R.nnz / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": base_config(desc, dst.ncells, src.ncells, K),
        "detail": {"nnz": R.nnz, "candidate_pairs": R.n_candidates},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port",
                         "sample": "full workload per step (restated reference algorithm: dual-DFS over bounding caps, "
                                   "Sutherland-Hodgman clip + area per pair, serial sparse(), serial mul! per level); Julia is "
                                   "not installed on this image"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "build_s": statistics.mean(t[0] for t in times),
        "apply_fwd_s": statistics.mean(t[1] for t in times), "apply_T_s": statistics.mean(t[2] for t in times),
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-halo", action="store_true", help="N > 1: build every block against the replicated source")
    ap.add_argument("--equal-blocks", dest="balanced_blocks", action="store_false",
                    help="N > 1: destination blocks of equal cell count instead of equal estimated candidate count "
                         "(default: balanced -- measured on config 5: 8 GPUs 1.79 -> 1.69 ms per step, 4 GPUs 2.00 -> 1.94)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from crg_b200 import _lib, grids
    from crg_b200.dist import ShardedRegridder, _LocalB200
    from crg_b200.regridder import Regridder, regrid_, transpose

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    hbm_peak, peak_src = load_peaks()
    counters, counters_state = kernel_counters()

    desc, fd, fs, K = WORKLOADS[args.workload]
    dst_spec, src_spec = fd(grids), fs(grids)
    n_dst, n_src = dst_spec.ncells, src_spec.ncells
    # all work runs on one non-default torch stream, handed to the library, so that torch CUDA
    # events bracket the library's kernels
    work_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(work_stream)
    stream = work_stream.cuda_stream
    rng = np.random.default_rng(20260101)
    x_host = make_field(n_src, K, rng)
    dims = 0 if K == 1 else 1                      # axis of the cells in a (K, n) C array = Julia dims=1 of (n, K)
    shp = lambda n: (n,) if K == 1 else (K, n)     # noqa: E731

    # device-resident inputs (for `value`): explicit vertices on one GPU; descriptors when sharded (every rank generates
    # its destination block and its source halo on the device inside the build)
    if world == 1:
        dst, src = dst_spec.materialize(), src_spec.materialize()
        dst_dev = grids.Grid(torch.from_numpy(dst.verts).to(dev), dst.manifold, None, dst.radius, dst.name, dst.meta)
        src_dev = grids.Grid(torch.from_numpy(src.verts).to(dev), src.manifold, None, src.radius, src.name, src.meta)
    x_dev = torch.from_numpy(x_host).to(dev)
    # 256 MiB buffer that is READ (summed) to evict the 126 MB L2 with clean lines before each apply
    # (a memset would leave the L2 full of dirty lines whose write-back is then charged to the apply)
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)

    tf = C.c_double()
    _lib.check(_lib.lib().crg_fp64_peak(-1, C.byref(tf)))
    fp64_peak = tf.value

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    state = {}

    def step_single(record):
        """Device-resident step on one GPU; returns per-phase event pairs when `record`."""
        e = [ev() for _ in range(6)] if record else None
        if record: e[0].record()
        R = Regridder(dst_dev, src_dev, stream=stream)
        if record: e[1].record()
        flush.sum()
        y = torch.empty(shp(n_dst), dtype=torch.float64, device=dev)
        if record: e[2].record()
        regrid_(y, R, x_dev, dims=dims, asynchronous=True)
        if record: e[3].record()
        flush.sum()
        xb = torch.empty(shp(n_src), dtype=torch.float64, device=dev)
        if record: e[4].record()
        regrid_(xb, transpose(R), y, dims=dims, asynchronous=True)
        if record: e[5].record()
        state["R"], state["y"], state["xb"] = R, y, xb
        return e

    sh_kw = dict(device=dev, halo=not args.no_halo)
    xs_dev = x_dev if K == 1 else x_dev.T.contiguous()      # sharded fields are (cells, K): blocks of cells are rows

    def step_sharded(record):
        e = [ev() for _ in range(6)] if record else None
        if record: e[0].record()
        factory = lambda rg, cg: _LocalB200(rg, cg, stream=stream)  # noqa: E731
        S = ShardedRegridder(dst_spec, src_spec, local_factory=factory, bounds=state.get("bounds"), **sh_kw)
        if record: e[1].record()
        flush.sum()
        if record: e[2].record()
        y = S.regrid(xs_dev if rank == 0 else None, trailing=() if K == 1 else (K,))   # NCCL broadcast + all-gather
        if record: e[3].record()
        flush.sum()
        if record: e[4].record()
        xb = S.regrid(y, transpose=True)                            # local A_r^T y_r + NCCL all-gather, overlap-add
        if record: e[5].record()
        state["R"], state["y"], state["xb"] = S, y, xb
        return e

    step = step_single if world == 1 else step_sharded
    if world > 1 and args.balanced_blocks:
        from crg_b200.dist import balanced_bounds, candidate_weights
        edges = torch.zeros(world + 1, dtype=torch.int64, device=dev)
        if rank == 0:
            dgrid = dst_spec.materialize()
            b = balanced_bounds(candidate_weights(dgrid, src_spec), world)
            edges = torch.tensor([0] + [hi for _, hi in b], dtype=torch.int64, device=dev)
        dist.broadcast(edges, src=0)
        e_ = edges.cpu().tolist()
        state["bounds"] = [(int(e_[k]), int(e_[k + 1])) for k in range(world)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step(False)
    barrier()
    launches0 = _lib.launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    start, stop = ev(), ev()
    start.record()
    events = [step(True) for _ in range(args.steps)]
    stop.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall0 + t_wall) if rank == 0 else None
    total_ms = start.elapsed_time(stop)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    R = state["R"]
    nnz = R.nnz if world > 1 else R.intersections.nnz
    value = nnz / (ms_per_step * 1e-3)

    build_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in events)
    fwd_ms = statistics.mean(e[2].elapsed_time(e[3]) for e in events)
    bwd_ms = statistics.mean(e[4].elapsed_time(e[5]) for e in events)

    # apply-only sequence (N = 1): alternating forward / transpose launches back to back, one event pair around the
    # whole sequence, i.e. without the ~5 us per-launch event/launch overhead that the in-step brackets include
    seq_pair_ms = None
    if world == 1:
        Rm, RmT = state["R"], transpose(state["R"])
        ys, xs_ = state["y"], state["xb"]
        for _ in range(3):
            regrid_(ys, Rm, x_dev, dims=dims, asynchronous=True); regrid_(xs_, RmT, ys, dims=dims, asynchronous=True)
        torch.cuda.synchronize()
        s0, s1 = ev(), ev()
        nseq = 20
        s0.record()
        for _ in range(nseq):
            regrid_(ys, Rm, x_dev, dims=dims, asynchronous=True); regrid_(xs_, RmT, ys, dims=dims, asynchronous=True)
        s1.record()
        torch.cuda.synchronize()
        seq_pair_ms = s0.elapsed_time(s1) / nseq

    # correctness guard on the timed result: conservation of the global mean (exact tilings only: the octahedral
    # stand-in of config 4 is not one, tests/test_gpu_parity_full.py)
    y, xb = state["y"], state["xb"]
    if world == 1:
        da, sa = torch.from_numpy(R.dst_areas.copy()).to(dev), torch.from_numpy(R.src_areas.copy()).to(dev)
        stats = R.intersections.stats()
        yv, xv, xbv = (y, x_dev, xb) if K == 1 else (y.T, x_dev.T, xb.T)
    else:
        da, sa = R.dst_areas, R.src_areas
        stats = R.local.stats
        yv, xv, xbv = y, xs_dev, xb
    bc = (lambda a, v: a if v.dim() == 1 else a[:, None])  # noqa: E731
    cons = abs(float((yv * bc(da, yv)).sum() / (xv * bc(sa, xv)).sum()) - 1.0) if rank == 0 or world == 1 else 0.0
    cons_T = abs(float((xbv * bc(sa, xbv)).sum() / (yv * bc(da, yv)).sum()) - 1.0)
    cons_tol = 1e-4 if args.workload == "cfg4" else 1e-11
    assert cons < cons_tol and cons_T < cons_tol, (cons, cons_T)

    parity = {"conservation_error": max(cons, cons_T)}
    collective_ms = None
    if world > 1:
        # N > 1: the gathered fields against a single-GPU regridder built here on rank 0 (per entry, not just conservation)
        if rank == 0:
            R1 = Regridder(dst_spec, src_spec, stream=stream)
            y1 = torch.empty_like(y); xb1 = torch.empty_like(xb)
            if K == 1:
                regrid_(y1, R1, xs_dev); regrid_(xb1, transpose(R1), y1)
            else:
                regrid_(y1, R1, xs_dev, dims=0); regrid_(xb1, transpose(R1), y1, dims=0)
            parity.update({"against": "single-GPU regridder on rank 0",
                           "max_rel_regrid_fwd": float(((y - y1).abs() / y1.abs().clamp_min(1e-300)).max()),
                           "max_rel_regrid_T": float(((xb - xb1).abs() / xb1.abs().clamp_min(1e-300)).max()),
                           "nnz_sharded": nnz, "nnz_single": R1.intersections.nnz})
            assert parity["max_rel_regrid_fwd"] < 1e-12 and parity["max_rel_regrid_T"] < 1e-11, parity
            del R1
        # per-collective device time of one step (separate, untimed pass)
        S = state["R"]
        lo, hi = S.dst_bounds[rank]
        tr = () if K == 1 else (K,)
        cm = {}
        for name, fn in (("broadcast_src_field", lambda: S._broadcast(xs_dev if rank == 0 else None, n_src, tr)),
                         ("all_gather_dst_blocks", lambda: S._all_gather_blocks(y[lo:hi].contiguous(), S.dst_bounds)),
                         ("transpose_all_gather_overlap_add", None)):
            barrier()
            a, b = ev(), ev()
            a.record()
            if fn is not None:
                for _ in range(5):
                    fn()
            else:
                for _ in range(5):
                    S.regrid(y, transpose=True)
            b.record()
            torch.cuda.synchronize()
            tt = torch.tensor([a.elapsed_time(b) / 5], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            cm[name] = float(tt.item())
        a, b = ev(), ev()
        part = torch.zeros((S.src_range[1] - S.src_range[0],) + tr, dtype=torch.float64, device=dev)
        barrier(); a.record()
        for _ in range(5):
            S.local.apply_T(part, y[lo:hi].contiguous(), True)
        b.record(); torch.cuda.synchronize()
        cm["transpose_all_gather_overlap_add"] = max(cm["transpose_all_gather_overlap_add"] - a.elapsed_time(b) / 5, 0.0)
        cm["dominant"] = max((k for k in cm), key=lambda k: cm[k])
        cm["bytes"] = {"broadcast_src_field": 8 * n_src * K, "all_gather_dst_blocks": 8 * n_dst * K,
                       "transpose_all_gather_overlap_add": 8 * K * sum(b_ - a_ for a_, b_ in S.halo_ranges())}
        collective_ms = cm

    line = None
    if rank == 0:
        # ---- rooflines (single-GPU kernels; for N > 1 they describe rank 0's shard) -------------
        n_cand = stats["n_candidates"]
        clip_ms = stats["ms_clip"]
        cc = counters["clip"]
        flops_per_pair = float(cc["flops_per_pair"]) if cc else CLIP_FLOPS_PER_PAIR_R01
        clip_tflops = n_cand * flops_per_pair / (clip_ms * 1e-3) / 1e12 if clip_ms > 0 else 0.0
        roof_clip = {"kernel": "clip_quad_kernel<3,128>", "bound": "fp64", "achieved": clip_tflops, "peak": fp64_peak,
                     "unit": "TFLOP/s", "frac": clip_tflops / fp64_peak if fp64_peak else None,
                     "traffic": (cc or {}).get("dram_bytes") if (world == 1 and args.workload == "cfg5") else None,
                     "peak_source": "crg_fp64_peak DFMA micro-benchmark, measured in this run (FP64 is not in MEASURED_PEAKS.json)",
                     "flops_per_pair": flops_per_pair,
                     "flops_per_pair_source": (f"ncu SASS counters of the shipped kernel on config 5 ({os.path.relpath(COUNTERS_JSON, ROOT)}, "
                                               f"source hash {cc['src_hash']})") if cc else
                                              f"counters {counters_state}: the r01 capture's figure is used",
                     "pairs": n_cand, "ms": clip_ms, "ms_is": "clip phase of the build (kernel + tile scan + compaction), CUDA events",
                     "pairs_per_s": n_cand / (clip_ms * 1e-3) if clip_ms > 0 else None,
                     "share_of_step": clip_ms / ms_per_step}
        # build as a whole against HBM: irreducible I/O and the traffic model of this
        # pipeline (pair list written + read, vertex gathers, COO written + read, P radix passes over 16 B records)
        n_loc_dst = n_dst if world == 1 else (R.dst_bounds[0][1] - R.dst_bounds[0][0])
        n_loc_src = n_src if world == 1 else (R.src_range[1] - R.src_range[0])
        nnz_loc = R.intersections.nnz if world == 1 else R.local.nnz
        passes = int(stats.get("sort_passes_csr", 0)) + int(stats.get("sort_passes_csc", 0))
        b_min = 96 * (n_loc_dst + n_loc_src) + 8 * (n_loc_dst + n_loc_src) + 2 * 12 * nnz_loc + 4 * (n_loc_dst + n_loc_src + 2)
        b_impl = b_min + n_cand * 8 * 2 + n_cand * 192 + nnz_loc * 16 * 2 + passes * 32 * nnz_loc
        build_dev_ms = stats["ms_device"]
        roof_build = {"bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "ms": build_dev_ms,
                      "bytes_impl_model": b_impl, "achieved": b_impl / (build_dev_ms * 1e-3) / 1e9,
                      "frac": b_impl / (build_dev_ms * 1e-3) / 1e9 / hbm_peak,
                      "bytes_min": b_min, "achieved_min": b_min / (build_dev_ms * 1e-3) / 1e9,
                      "radix_passes": passes, "candidate_pairs": n_cand, "nnz": nnz_loc,
                      "candidates_per_nnz": n_cand / nnz_loc if nnz_loc else None,
                      "note": "the build is bounded by the clip kernel (L1 / FP64), the radix passes and broad-phase latency, "
                              "not by HBM (DESIGN.md section 4)"}
        builds = sorted(e[0].elapsed_time(e[1]) for e in events)
        roof_apply = None
        if world == 1:
            by_f = R.intersections.apply_bytes(K, True)
            by_t = transpose(R).intersections.apply_bytes(K, True)
            apply_f_gbs = by_f / (fwd_ms * 1e-3) / 1e9
            apply_t_gbs = by_t / (bwd_ms * 1e-3) / 1e9
            cf, ct = counters["spmv_fwd"], counters["spmv_T"]
            kern = "spmv_sell_kernel (forward regrid!)" if K == 1 else "spmm_cf_kernel (forward regrid!, K levels, cell-fastest)"
            roof_apply = {"kernel": kern, "bound": "hbm", "achieved": apply_f_gbs,
                          "peak": hbm_peak, "unit": "GB/s", "frac": apply_f_gbs / hbm_peak,
                          "traffic": (cf or {}).get("dram_bytes") if (args.workload == "cfg5") else None,
                          "peak_source": peak_src, "bytes": by_f, "ms": fwd_ms,
                          "transpose": {"achieved": apply_t_gbs, "frac": apply_t_gbs / hbm_peak, "bytes": by_t, "ms": bwd_ms,
                                        "traffic": (ct or {}).get("dram_bytes") if (args.workload == "cfg5") else None},
                          "back_to_back_fwd_plus_transpose": {
                              "pair_ms": seq_pair_ms, "bytes": by_f + by_t,
                              "achieved": (by_f + by_t) / (seq_pair_ms * 1e-3) / 1e9,
                              "frac": (by_f + by_t) / (seq_pair_ms * 1e-3) / 1e9 / hbm_peak,
                              "note": "20 alternating launches between one event pair"}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": base_config(desc, n_dst, n_src, K),
            "detail": {"nnz": nnz, "candidate_pairs": n_cand,
                       "parallelism": ((f"dst-sharded x{world}, source halo per rank" if not args.no_halo else f"dst-sharded x{world}, replicated source")
                                       + (", blocks balanced by estimated candidate count" if args.balanced_blocks else ", blocks of equal cell count")) if world > 1 else "single GPU",
                       "inputs": "explicit cell vertices resident in HBM" if world == 1 else "described grids: every rank generates its destination block and source halo on the device",
                       "l2": "256 MiB buffer read before each apply (clean L2 eviction); build working set (>1 GB) exceeds the 126 MB L2"},
            "build_ms": build_ms, "apply_fwd_ms": fwd_ms, "apply_T_ms": bwd_ms,
            "build_phases_ms": {k[3:]: round(v, 4) for k, v in stats.items() if k.startswith("ms_")},
            "candidate_pairs_per_s": n_cand / (build_ms * 1e-3), "wall_s_timed_region": t_wall,
            # `roofline` = the dominant kernel of the step, the clip (FP64); the HBM-bound regrid! kernels the north
            # star sets its 70 % target on are in `roofline_apply`
            "roofline": roof_clip, "roofline_apply": roof_apply, "roofline_build": roof_build,
            "kernel_counters": counters_state,
            "build_ms_median": builds[len(builds) // 2], "build_ms_best": builds[0],
            "nnz_per_s_build": nnz_loc / (build_ms * 1e-3),
            "gpu_launches": launches, "clocks": clocks, "parity": parity,
        }
        if collective_ms is not None:
            line["collective_ms"] = collective_ms
            line["per_rank"] = {"n_dst_block": n_loc_dst, "n_src_halo": n_loc_src, "nnz_block": nnz_loc}

    # ---- e2e: host (pinned) buffers in and out, copies inside the timed region --------------------
    def pinned(a):
        t_ = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t_, t_.numpy()
    keep = []
    if world == 1:
        dv_t, dv = pinned(dst.verts); sv_t, sv = pinned(src.verts)
        xh_t, xh = pinned(x_host)
        yh_t, yh = pinned(np.zeros(shp(n_dst))); xbh_t, xbh = pinned(np.zeros(shp(n_src)))
        keep += [dv_t, sv_t, xh_t, yh_t, xbh_t]
        dst_h = grids.Grid(dv, dst.manifold); src_h = grids.Grid(sv, src.manifold)

        def step_e2e(explicit):
            # explicit: both vertex soups are uploaded; otherwise the grids are passed as the few numbers that
            # describe them and their cells are generated on the device inside the build
            t_ = [time.perf_counter()]
            R_ = Regridder(dst_h, src_h, stream=stream) if explicit else Regridder(dst_spec, src_spec, stream=stream)
            t_.append(time.perf_counter())
            regrid_(yh, R_, xh, dims=dims)                   # H2D x, D2H y
            t_.append(time.perf_counter())
            regrid_(xbh, transpose(R_), yh, dims=dims)       # H2D y, D2H xb
            t_.append(time.perf_counter())
            da_, sa_ = R_.dst_areas, R_.src_areas            # D2H of both area vectors (lazy otherwise)
            t_.append(time.perf_counter())
            for i_, k_ in enumerate(("build", "regrid_fwd", "regrid_T", "areas")):
                e2e_parts[k_] = e2e_parts.get(k_, 0.0) + (t_[i_ + 1] - t_[i_]) * 1e3
            return R_
        n_e2e = max(3, min(args.steps, 10))
        res = {}
        e2e_parts = {}
        for explicit in (True, False):
            for _ in range(2):
                step_e2e(explicit)
            torch.cuda.synchronize()
            e2e_parts.clear()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                R_ = step_e2e(explicit)
            torch.cuda.synchronize()
            res[explicit] = (1e3 * (time.perf_counter() - t0) / n_e2e, R_.intersections.stats(),
                             {k_: round(v_ / n_e2e, 3) for k_, v_ in e2e_parts.items()})
            assert np.allclose(yh, y.cpu().numpy(), rtol=1e-11)
        field_h2d = x_host.nbytes + yh.nbytes
        d2h = 8 * (n_dst + n_src) + yh.nbytes + xbh.nbytes
        e2e_ms, st_, parts_ = res[False]
        line["e2e"] = {"value": nnz / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(field_h2d),
                       "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "steps": n_e2e,
                       "inputs": "grids passed as descriptors (what the Julia glue passes for Healpix / lon-lat / RingGrids "
                                 "grid objects): cell vertices are generated on the device inside every build; fields from / "
                                 "to pinned host memory",
                       "host_ms": parts_,
                       "build_phases_ms": {k[3:]: round(v, 3) for k, v in st_.items() if k.startswith("ms_")}}
        e2e_ms, st_, parts_x = res[True]
        line["e2e_explicit_cells"] = {
            "value": nnz / (e2e_ms * 1e-3), "unit": UNIT,
            "h2d_bytes_per_step": int(dst.verts.nbytes + src.verts.nbytes + field_h2d),
            "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "steps": n_e2e,
            "inputs": "both grids as explicit vertex soups in pinned host memory (any grid type through Trees.getcell), "
                      "uploaded inside every build: PCIe-bound",
            "host_ms": parts_x,
            "build_phases_ms": {k[3:]: round(v, 3) for k, v in st_.items() if k.startswith("ms_")}}
    else:
        # N > 1: described grids; the source field comes from pinned host memory on rank 0, the results (both fields and
        # both area vectors) go back to the host on rank 0 only.
        xh_t, _ = pinned(x_host if K == 1 else np.ascontiguousarray(x_host.T))
        outs_h = [torch.empty(s_, dtype=torch.float64).pin_memory() for s_ in
                  ((n_dst,) + (() if K == 1 else (K,)), (n_src,) + (() if K == 1 else (K,)), (n_dst,), (n_src,))] if rank == 0 else []
        keep += [xh_t] + outs_h

        def step_e2e():
            factory = lambda rg, cg: _LocalB200(rg, cg, stream=stream)  # noqa: E731
            S = ShardedRegridder(dst_spec, src_spec, local_factory=factory, bounds=state.get("bounds"), **sh_kw)
            xd = xh_t.to(dev, non_blocking=True) if rank == 0 else None
            y_ = S.regrid(xd, trailing=() if K == 1 else (K,))
            xb_ = S.regrid(y_, transpose=True)
            da_, sa_ = S.dst_areas, S.src_areas                      # collective
            if rank == 0:
                for h_, d_ in zip(outs_h, (y_, xb_, da_, sa_)):
                    h_.copy_(d_, non_blocking=True)
            torch.cuda.synchronize()
        step_e2e()
        barrier()
        n_e2e = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            step_e2e()
        barrier()
        e2e_ms = 1e3 * (time.perf_counter() - t0) / n_e2e
        tt = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
        if rank == 0:
            line["e2e"] = {"value": nnz / (e2e_ms * 1e-3), "unit": UNIT,
                           "h2d_bytes_per_step": int(x_host.nbytes),
                           "d2h_bytes_per_step": int(8 * (K + 1) * (n_dst + n_src)), "ms_per_step": e2e_ms, "steps": n_e2e,
                           "inputs": "grids passed as descriptors: every rank generates its destination block and source halo on "
                                     "its device; source field from pinned host memory on rank 0; both fields and both area "
                                     "vectors read back on rank 0"}

    # ---- cpu baseline beside it (rank 0, N = 1 only) -------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        from oracle.parity import parity_report
        oracle.build()
        nthreads = oracle.use_all_cores()
        trees = oracle_trees(oracle, dst, src)
        t0 = time.perf_counter()
        Rc, yc, tcpu = cpu_reference_step(oracle, dst, src, trees, x_host, nthreads)
        cpu_s = time.perf_counter() - t0
        line["cpu_baseline"] = {
            "value": Rc.nnz / cpu_s, "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": f"1 full step of the same workload ({cpu_s:.1f} s: build {tcpu[0]:.2f} s, regrid! fwd "
                      f"{tcpu[1]*1e3:.1f} ms, transpose {tcpu[2]*1e3:.1f} ms); restated reference algorithm "
                      "(oracle/), tree construction excluded; Julia not installed",
            "build_s": tcpu[0], "apply_fwd_s": tcpu[1], "apply_T_s": tcpu[2]}
        # parity of the timed GPU result against the CPU restatement (the checker, same data): matrix entry by entry
        # in the terms of north_star + the regridded field
        rep = parity_report(R.intersections.tocsc(), Rc.tocsc(), Rc.dst_areas, Rc.src_areas)
        parity.update({"against": "CPU restatement of the reference (oracle/), same grids",
                       "max_rel": rep["max_rel_above_floor"], "max_abs": rep["max_abs"],
                       "n_pattern_diff_above_tau": rep["n_pattern_diff_above_tau"], "tau": rep["tau"],
                       "n_entries_beyond_1e-10_rel": rep["n_entries_beyond_tolerance"],
                       "nnz_device": rep["nnz_device"], "nnz_oracle": rep["nnz_oracle"],
                       "n_under_tau_device": rep["n_under_tau_device"], "n_under_tau_oracle": rep["n_under_tau_oracle"],
                       "symdiff_max_value": rep["symdiff_max_value"],
                       "max_rel_diff_regrid": float(np.max(np.abs(y.cpu().numpy() / yc - 1)))})
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# FP64 flops per candidate pair of clip_quad_kernel<3,128> on the cfg5 workload in the r01 capture (111.1 DFMA x 2 +
# 86.5 DMUL + 30.7 DADD per pair) -- only used when profiles/r02_kernel_counters.json is missing or stale.
CLIP_FLOPS_PER_PAIR_R01 = 339.4

if __name__ == "__main__":
    main()
