#!/usr/bin/env python
"""bench.py -- headline benchmark of the two hot paths on B200 (see DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg5|cfg2|cfg1]

One *step* = one pass of the hot path over the workload: build the Regridder (broad phase ->
clip/area -> sort/assemble CSR + CSC -> areas), then regrid! forward and regrid! with
transpose(R).  Default workload = BASELINE.json configs[4]: 0.25 deg lon-lat (1440 x 720,
destination) <-> HEALPix nside=512 ring (source); it fits one GPU.  For N > 1 the destination
cells (forward) and the source cells (transpose) are sharded over the ranks ("strong" scaling:
total work fixed), fields are broadcast / all-gathered with NCCL.

metric = overlapping cell pairs (= nnz of the regridder, identical for every implementation)
processed per second through the whole step.  `value`: vertices and fields already resident in
HBM; `e2e`: host (pinned) buffers in, host buffers out, copies inside the timed region.
`--impl reference` times the CPU restatement of the reference algorithm (oracle/, OpenMP on all
host cores) on the same workload: Julia is not installed here, so this is kind="port".
"""
from __future__ import annotations

import argparse
import ctypes as C
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

WORKLOADS = {
    # name: (description, dst spec factory, src spec factory)   (GridSpec.materialize() gives the explicit cells)
    "cfg5": ("0.25deg lon-lat 1440x720 (dst) <-> HEALPix nside=512 ring (src): build + regrid! fwd + transpose",
             lambda g: g.lonlat_spec(1440, 720), lambda g: g.healpix_spec(512, "ring")),
    "cfg2": ("0.5deg lon-lat 720x360 (dst) <-> HEALPix nside=256 ring (src): build + regrid! fwd + transpose",
             lambda g: g.lonlat_spec(720, 360), lambda g: g.healpix_spec(256, "ring")),
    "cfg1": ("2deg lon-lat 180x90 (dst) <- 1deg lon-lat 360x180 (src): build + regrid! fwd + transpose",
             lambda g: g.lonlat_spec(180, 90), lambda g: g.lonlat_spec(360, 180)),
}


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


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the restated reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------

def cpu_reference_step(oracle, dst, src, trees, x, nthreads):
    """One step on the CPU: dual-DFS candidates (threaded) -> per-pair clip + area (threaded) ->
    serial sparse() -> serial areas -> serial mul! forward and transposed (the reference's own
    threading model).  Tree construction is excluded (the reference's trees
    are lazy, O(1))."""
    t0 = time.perf_counter()
    cands = oracle.dual_query(trees[1], trees[0], nthreads)
    R = oracle.build_regridder(dst, src, candidates=cands, nthreads=nthreads)
    t1 = time.perf_counter()
    y = R.regrid(x)
    t2 = time.perf_counter()
    R.regrid(y, transpose=True)
    t3 = time.perf_counter()
    return R, (t1 - t0, t2 - t1, t3 - t2)


def run_reference(args, rank):
    if rank != 0:
        return
    from crg_b200 import grids
    from oracle import oracle
    oracle.build()
    desc, fd, fs = WORKLOADS[args.workload]
    dst, src = fd(grids).materialize(), fs(grids).materialize()
    nthreads = oracle.use_all_cores()          # torchrun exports OMP_NUM_THREADS=1
    trees = (oracle.treeify(dst), oracle.treeify(src))
    x = np.random.default_rng(20260101).random(src.ncells)
    times = []
    R = None
    for it in range(args.warmup + args.steps):
        R, t = cpu_reference_step(oracle, dst, src, trees, x, nthreads)
        if it >= args.warmup:
            times.append(t)
    tot = sum(sum(t) for t in times)
    ms = 1e3 * tot / len(times)
    value = R.nnz / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "n_dst": dst.ncells, "n_src": src.ncells, "nnz": R.nnz,
                   "candidate_pairs": R.n_candidates},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port",
                         "sample": "full workload per step (restated reference algorithm: dual-DFS over bounding caps, "
                                   "Sutherland-Hodgman clip + area per pair, serial sparse(), serial mul!); Julia is not "
                                   "installed on this image"},
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
    ap.add_argument("--balanced-blocks", action="store_true",
                    help="N > 1: destination blocks of equal estimated candidate count instead of equal cell count "
                         "(measured on cfg5, 4 GPUs: 2.281 vs 2.285 ms per step, e2e 5.8 vs 4.8 ms -- fixed per-rank "
                         "costs dominate and uneven blocks pay a padded all-gather, so the default stays equal counts)")
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
    from crg_b200.dist import ShardedRegridder, _LocalB200, block_bounds
    from crg_b200.regridder import Regridder, regrid_, transpose

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    hbm_peak, peak_src = load_peaks()

    desc, fd, fs = WORKLOADS[args.workload]
    dst_spec, src_spec = fd(grids), fs(grids)
    dst, src = dst_spec.materialize(), src_spec.materialize()
    n_dst, n_src = dst.ncells, src.ncells
    # all work runs on one non-default torch stream, handed to the library, so that torch CUDA
    # events bracket the library's kernels (the legacy default stream's handle is 0 == "own stream")
    work_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(work_stream)
    stream = work_stream.cuda_stream
    rng = np.random.default_rng(20260101)
    x_host = rng.random(n_src)

    # device-resident inputs (for `value`)
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
        y = torch.empty(n_dst, dtype=torch.float64, device=dev)
        if record: e[2].record()
        regrid_(y, R, x_dev, asynchronous=True)
        if record: e[3].record()
        flush.sum()
        xb = torch.empty(n_src, dtype=torch.float64, device=dev)
        if record: e[4].record()
        regrid_(xb, transpose(R), y, asynchronous=True)
        if record: e[5].record()
        state["R"], state["y"], state["xb"] = R, y, xb
        return e

    def step_sharded(record):
        e = [ev() for _ in range(6)] if record else None
        if record: e[0].record()
        factory = lambda rg, cg: _LocalB200(rg, cg, stream=stream)  # noqa: E731
        S = ShardedRegridder(dst_dev, src_dev, local_factory=factory, device=dev, bounds=state.get("bounds"))
        if record: e[1].record()
        flush.sum()
        if record: e[2].record()
        y = S.regrid(x_dev if rank == 0 else None)                  # NCCL broadcast + all-gather
        if record: e[3].record()
        flush.sum()
        if record: e[4].record()
        xb = S.regrid(y, transpose=True)                            # local A_r^T y_r + NCCL all-reduce
        if record: e[5].record()
        state["R"], state["y"], state["xb"] = S, y, xb
        return e

    step = step_single if world == 1 else step_sharded
    if world > 1 and args.balanced_blocks:
        # destination blocks of equal estimated candidate count instead of equal cell count, decided once
        # (a property of the two grids, like the partition of any sharded operator) and reused by every build
        from crg_b200.dist import balanced_bounds, candidate_weights
        edges = torch.zeros(world + 1, dtype=torch.int64, device=dev)
        if rank == 0:
            b = balanced_bounds(candidate_weights(dst_dev, src_dev), world)
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

    # apply-only sequence (N = 1): alternating forward / transpose launches back to back -- two different
    # 111 MB matrices + vectors (340 MB per pair > 126 MB L2) -- one event pair around the whole sequence,
    # i.e. without the ~5 us per-launch event/launch overhead that the in-step brackets include
    seq_pair_ms = None
    if world == 1:
        Rm, RmT = state["R"], transpose(state["R"])
        ys, xs_ = state["y"], state["xb"]
        for _ in range(3):
            regrid_(ys, Rm, x_dev, asynchronous=True); regrid_(xs_, RmT, ys, asynchronous=True)
        torch.cuda.synchronize()
        s0, s1 = ev(), ev()
        nseq = 20
        s0.record()
        for _ in range(nseq):
            regrid_(ys, Rm, x_dev, asynchronous=True); regrid_(xs_, RmT, ys, asynchronous=True)
        s1.record()
        torch.cuda.synchronize()
        seq_pair_ms = s0.elapsed_time(s1) / nseq

    # correctness guard on the timed result: conservation of the global mean
    y, xb = state["y"], state["xb"]
    if world == 1:
        da, sa = torch.from_numpy(R.dst_areas).to(dev), torch.from_numpy(R.src_areas).to(dev)
        stats = R.intersections.stats()
    else:
        da, sa = R.dst_areas, R.src_areas
        stats = R.local.stats
    cons = abs(float((y * da).sum() / (x_dev * sa).sum()) - 1.0) if rank == 0 or world == 1 else 0.0
    cons_T = abs(float((xb * sa).sum() / (y * da).sum()) - 1.0)
    assert cons < 1e-11 and cons_T < 1e-11, (cons, cons_T)

    line = None
    if rank == 0:
        # ---- rooflines (single-GPU kernels; for N > 1 they describe rank 0's shard) -------------
        n_cand = stats["n_candidates"]
        clip_ms = stats["ms_clip"]
        # K3: measured FP64 work per pair from the ncu capture in profiles/ 
        flops_per_pair = float(os.environ.get("CRG_CLIP_FLOPS_PER_PAIR", "0") or 0) or CLIP_FLOPS_PER_PAIR
        clip_tflops = n_cand * flops_per_pair / (clip_ms * 1e-3) / 1e12 if clip_ms > 0 else 0.0
        if world == 1:
            by_f = R.intersections.apply_bytes(1, True)
            by_t = transpose(R).intersections.apply_bytes(1, True)
        else:
            lo, hi = R.dst_bounds[0]
            by_f = 12 * R.local.nnz + 4 * (hi - lo + 1) + 8 * (hi - lo) + 8 * (n_src + hi - lo)
            by_t = None
        apply_f_gbs = by_f / (fwd_ms * 1e-3) / 1e9 if world == 1 else None
        apply_t_gbs = by_t / (bwd_ms * 1e-3) / 1e9 if world == 1 else None
        roof_clip = {"kernel": "clip_quad_kernel<3,128>", "bound": "fp64", "achieved": clip_tflops, "peak": fp64_peak,
                     "unit": "TFLOP/s", "frac": clip_tflops / fp64_peak if fp64_peak else None, "traffic": None,
                     "peak_source": "crg_fp64_peak DFMA micro-benchmark, measured in this run",
                     "flops_per_pair": flops_per_pair, "pairs": n_cand, "ms": clip_ms,
                     "pairs_per_s": n_cand / (clip_ms * 1e-3) if clip_ms > 0 else None}
        # build as a whole against HBM: irreducible I/O and the traffic model of this
        # pipeline (pair list written + read, vertex gathers, COO written + read, P radix passes over 16 B records)
        n_loc_dst = n_dst if world == 1 else (R.dst_bounds[0][1] - R.dst_bounds[0][0])
        nnz_loc = R.intersections.nnz if world == 1 else R.local.nnz
        passes = int(stats.get("sort_passes_csr", 0)) + int(stats.get("sort_passes_csc", 0))
        b_min = 96 * (n_loc_dst + n_src) + 8 * (n_loc_dst + n_src) + 2 * 12 * nnz_loc + 4 * (n_loc_dst + n_src + 2)
        b_impl = b_min + n_cand * 8 * 2 + n_cand * 192 + nnz_loc * 16 * 2 + passes * 32 * nnz_loc
        build_dev_ms = stats["ms_device"]
        roof_build = {"bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "ms": build_dev_ms,
                      "bytes_impl_model": b_impl, "achieved": b_impl / (build_dev_ms * 1e-3) / 1e9,
                      "frac": b_impl / (build_dev_ms * 1e-3) / 1e9 / hbm_peak,
                      "bytes_min": b_min, "achieved_min": b_min / (build_dev_ms * 1e-3) / 1e9,
                      "radix_passes": passes, "candidate_pairs": n_cand, "nnz": nnz_loc,
                      "candidates_per_nnz": n_cand / nnz_loc if nnz_loc else None,
                      "note": "the build is bounded by the FP64 clip, the radix passes and broad-phase latency, "
                              "not by HBM (DESIGN.md section 4)"}
        builds = sorted(e[0].elapsed_time(e[1]) for e in events)
        roof_apply = None
        if world == 1:
            roof_apply = {"kernel": "spmv_sell_kernel<true> (forward regrid!)", "bound": "hbm", "achieved": apply_f_gbs,
                          "peak": hbm_peak, "unit": "GB/s", "frac": apply_f_gbs / hbm_peak, "traffic": SPMV_TRAFFIC_BYTES,
                          "peak_source": peak_src, "bytes": by_f, "ms": fwd_ms,
                          "transpose": {"achieved": apply_t_gbs, "frac": apply_t_gbs / hbm_peak, "bytes": by_t, "ms": bwd_ms},
                          "back_to_back_fwd_plus_transpose": {
                              "pair_ms": seq_pair_ms, "bytes": by_f + by_t,
                              "achieved": (by_f + by_t) / (seq_pair_ms * 1e-3) / 1e9,
                              "frac": (by_f + by_t) / (seq_pair_ms * 1e-3) / 1e9 / hbm_peak,
                              "note": "20 alternating launches between one event pair; inputs (340 MB per pair) exceed L2"}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "n_dst": n_dst, "n_src": n_src, "nnz": nnz, "candidate_pairs": n_cand,
                       "parallelism": f"dst-sharded x{world}" if world > 1 else "single GPU",
                       "l2": "256 MiB buffer read before each apply (clean L2 eviction); build working set (>1 GB) exceeds the 126 MB L2"},
            "build_ms": build_ms, "apply_fwd_ms": fwd_ms, "apply_T_ms": bwd_ms,
            "build_phases_ms": {k[3:]: round(v, 4) for k, v in stats.items() if k.startswith("ms_")},
            "candidate_pairs_per_s": n_cand / (build_ms * 1e-3), "wall_s_timed_region": t_wall,
            # `roofline` = the HBM-bound regrid! kernel the north star sets its target on; the kernel with the
            # largest share of the step is the FP64-bound clip kernel, reported beside it in `roofline_clip`
            "roofline": roof_apply if roof_apply else roof_clip,
            "roofline_clip": roof_clip, "roofline_apply": roof_apply, "roofline_build": roof_build,
            "build_ms_median": builds[len(builds) // 2], "build_ms_best": builds[0],
            "nnz_per_s_build": nnz_loc / (build_ms * 1e-3),
            "dominant_kernel_by_time": "clip_quad_kernel (FP64-bound, see roofline_clip): %.0f%% of the step" % (100 * clip_ms / ms_per_step),
            "gpu_launches": launches, "clocks": clocks,
            "conservation_error": max(cons, cons_T),
        }

    # ---- e2e: host (pinned) buffers in and out, copies inside the timed region --------------------
    def pinned(a):
        t_ = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t_, t_.numpy()
    keep = []
    if world == 1:
        dv_t, dv = pinned(dst.verts); sv_t, sv = pinned(src.verts)
        xh_t, xh = pinned(x_host)
        yh_t, yh = pinned(np.zeros(n_dst)); xbh_t, xbh = pinned(np.zeros(n_src))
        keep += [dv_t, sv_t, xh_t, yh_t, xbh_t]
        dst_h = grids.Grid(dv, dst.manifold); src_h = grids.Grid(sv, src.manifold)

        def step_e2e(explicit):
            # explicit: both vertex soups are uploaded (435 MB); otherwise the grids are passed as the few
            # numbers that describe them and their cells are generated on the device inside the build
            t_ = [time.perf_counter()]
            R_ = Regridder(dst_h, src_h, stream=stream) if explicit else Regridder(dst_spec, src_spec, stream=stream)
            t_.append(time.perf_counter())
            regrid_(yh, R_, xh)                              # H2D x, D2H y
            t_.append(time.perf_counter())
            regrid_(xbh, transpose(R_), yh)                  # H2D y, D2H xb
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
        if rank == 0:
            e2e_ms, st_, parts_ = res[False]
            line["e2e"] = {"value": nnz / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(field_h2d),
                           "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "steps": n_e2e,
                           "inputs": "grids passed as descriptors (lon-lat 1440x720, HEALPix nside 512): cell vertices are "
                                     "generated on the device inside every build; fields from / to pinned host memory",
                           "host_ms": parts_,
                           "build_phases_ms": {k[3:]: round(v, 3) for k, v in st_.items() if k.startswith("ms_")}}
            e2e_ms, st_, parts_x = res[True]
            line["e2e_explicit_cells"] = {
                "value": nnz / (e2e_ms * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": int(dst.verts.nbytes + src.verts.nbytes + field_h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "steps": n_e2e,
                "inputs": "both grids as explicit vertex soups in pinned host memory, uploaded inside every build",
                "host_ms": parts_x,
                "build_phases_ms": {k[3:]: round(v, 3) for k, v in st_.items() if k.startswith("ms_")}}
    else:
        # N > 1: the grids are passed as descriptors; every rank generates the destination cells on its
        # device (and keeps its block), the replicated source cells are generated inside the local build;
        # the source field comes from pinned host memory on rank 0, the results go back to the host.
        from crg_b200.regridder import grid_cells
        xh_t, xh = pinned(x_host)
        outs_h = [torch.empty(n, dtype=torch.float64).pin_memory() for n in (n_dst, n_src, n_dst, n_src)]
        keep += [xh_t] + outs_h

        def step_e2e():
            dt = torch.empty((n_dst, 4, 3), dtype=torch.float64, device=dev)
            grid_cells(dst_spec, out=dt)
            factory = lambda rg, cg: _LocalB200(rg, cg, stream=stream)  # noqa: E731
            S = ShardedRegridder(grids.Grid(dt, dst.manifold), src_spec, local_factory=factory, device=dev, bounds=state.get("bounds"))
            xd = xh_t.to(dev, non_blocking=True) if rank == 0 else None
            y_ = S.regrid(xd)
            xb_ = S.regrid(y_, transpose=True)
            for h_, d_ in zip(outs_h, (y_, xb_, S.dst_areas, S.src_areas)):
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
                           "d2h_bytes_per_step": int(2 * 8 * (n_dst + n_src)), "ms_per_step": e2e_ms, "steps": n_e2e,
                           "inputs": "grids passed as descriptors, cells generated on every rank's device; source field "
                                     "from pinned host memory on rank 0; both fields and both area vectors read back"}

    # ---- cpu baseline beside it (rank 0, N = 1 only) -------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        oracle.build()
        nthreads = oracle.use_all_cores()
        trees = (oracle.treeify(dst), oracle.treeify(src))
        t0 = time.perf_counter()
        Rc, tcpu = cpu_reference_step(oracle, dst, src, trees, x_host, nthreads)
        cpu_s = time.perf_counter() - t0
        line["cpu_baseline"] = {
            "value": Rc.nnz / cpu_s, "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": f"1 full step of the same workload ({cpu_s:.1f} s: build {tcpu[0]:.2f} s, regrid! fwd "
                      f"{tcpu[1]*1e3:.1f} ms, transpose {tcpu[2]*1e3:.1f} ms); restated reference algorithm "
                      "(oracle/), tree construction excluded; Julia not installed",
            "build_s": tcpu[0], "apply_fwd_s": tcpu[1], "apply_T_s": tcpu[2]}
        # parity of the timed GPU result against the CPU baseline's matrix (cheap, same data)
        line["cpu_baseline"]["max_rel_diff_regrid"] = float(np.max(np.abs(y.cpu().numpy() / Rc.regrid(x_host) - 1)))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# FP64 flops per candidate pair of clip_quad_kernel<3,128> on the cfg5 workload, counted from the SASS page
# of the ncu capture summarised in profiles/README.md (111.1 DFMA x 2 + 86.5 DMUL + 30.7 DADD per pair;
# the first, division-based kernel of this round executed 339.6).
CLIP_FLOPS_PER_PAIR = 339.4
# dram__bytes_read.sum + dram__bytes_write.sum of one forward spmv launch on cfg5 (ncu --set full)
SPMV_TRAFFIC_BYTES = 157804800

if __name__ == "__main__":
    main()
