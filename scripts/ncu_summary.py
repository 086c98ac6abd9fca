"""Summarise an .ncu-rep (raw page) into the handful of counters DESIGN.md / profiles quote.

  python scripts/ncu_summary.py rep.ncu-rep                     human-readable summary of every kernel in the report
  python scripts/ncu_summary.py --json clip=rep[:pairs] spmv_fwd=rep spmv_T=rep
        writes profiles/r02_kernel_counters.json: the measured FP64 work per candidate pair of the clip kernel
        (SASS thread-instruction counters) and the DRAM traffic per launch of the clip / SpMV kernels, each with the
        hash of the kernel sources it was measured on -- bench.py reads its roofline constants from there and
        refuses them when the sources have changed since (tests/test_host.py checks the same)."""
import csv, hashlib, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "conservativeregridding.jl_b200", "csrc")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__lsu_writeback_active_mem_lgds.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max"]
SRC = {"clip": ["geom.cuh", "kernels.cuh"], "spmv_fwd": ["sell.cuh"], "spmv_T": ["sell.cuh"]}
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0,
        "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}

def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]

def main(path):
    hdr, units, rows = raw(path)
    for vals in rows:
        print("## kernel:", vals[hdr.index("Kernel Name")][:110], " grid", vals[hdr.index("Grid Size")], "block", vals[hdr.index("Block Size")])
        for i, h in enumerate(hdr):
            if h in KEYS or ("issue_stalled" in h and "per_issue_active" in h and vals[i] and float(vals[i]) >= 0.3):
                print(f"  {h} [{units[i]}] = {vals[i]}")

def sass_thread_ops(path, kernel_substr):
    """Predicated-on THREAD instruction counts by opcode from the SASS page of one kernel of the report."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    ops, cur, hdr, seen = {}, None, None, 0
    for r in rows:
        if len(r) >= 2 and r[0] in ("Function Name", "Kernel Name"):
            cur = r[1]
            seen += kernel_substr in cur          # (the page lists every kernel once per view: count the first listing only)
            if seen > 1 and kernel_substr in cur: cur = None
            continue
        if r and r[0] == "Address": hdr = r; continue
        if hdr and cur and kernel_substr in cur and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            src = d.get("Source", "").strip().split()
            if not src: continue
            op = src[1] if src[0].startswith("@") and len(src) > 1 else src[0]
            op = op.split(".")[0]
            try: n = int(d.get("Predicated-On Thread Instructions Executed") or 0)
            except ValueError: continue
            ops[op] = ops.get(op, 0) + n
    return ops

def src_hash(files):
    h = hashlib.sha256()
    for f in files:
        h.update(open(os.path.join(CSRC, f), "rb").read())
    return h.hexdigest()[:16]

def to_json(specs):
    out = {}
    for spec in specs:
        name, _, rest = spec.partition("=")
        path, _, pairs = rest.partition(":")
        hdr, units, rows = raw(path)
        pick = {"clip": "clip_quad_kernel", "spmv_fwd": "spmv_sell_kernel", "spmv_T": "spmv_sell_pipelined_kernel"}[name]
        vals = [v for v in rows if pick in v[hdr.index("Kernel Name")]][-1]
        def get(key):
            i = hdr.index(key)
            return float(vals[i].replace(",", "")) * UNIT.get(units[i], 1.0)
        e = {"kernel": vals[hdr.index("Kernel Name")][:80], "report": os.path.basename(path),
             "dram_bytes": get("dram__bytes_read.sum") + get("dram__bytes_write.sum"),
             "duration_us": get("gpu__time_duration.sum") * 1e6, "src_files": SRC[name], "src_hash": src_hash(SRC[name])}
        if name == "clip":
            ops = sass_thread_ops(path, pick)
            dfma, dmul, dadd = (float(ops.get(k, 0)) for k in ("DFMA", "DMUL", "DADD"))
            n = float(pairs)
            e["dsetp_per_pair"] = ops.get("DSETP", 0) / n
            e["thread_instructions_per_pair"] = sum(ops.values()) / n
            e.update({"pairs": n, "dfma_per_pair": dfma / n, "dmul_per_pair": dmul / n, "dadd_per_pair": dadd / n,
                      "flops_per_pair": (2 * dfma + dmul + dadd) / n,
                      "fp64_pipe_active_pct": get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                      "l1_data_pipe_wavefronts_pct": get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                      "lanes_per_instruction": get("smsp__thread_inst_executed_per_inst_executed.ratio")})
        out[name] = e
    p = os.path.join(ROOT, "profiles", "r02_kernel_counters.json")
    json.dump(out, open(p, "w"), indent=1)
    print(open(p).read())

if __name__ == "__main__":
    if sys.argv[1] == "--json": to_json(sys.argv[2:])
    else: main(sys.argv[1])
