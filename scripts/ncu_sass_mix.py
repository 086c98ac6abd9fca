"""Instruction mix of one kernel of an .ncu-rep (needs --import-source on, -lineinfo): executed WARP instructions by
opcode and by source line, split into FP64-pipe opcodes and the rest.  usage: ncu_sass_mix.py rep kernel_substr [topN]"""
import csv, io, subprocess, sys, collections
FP64 = {"DFMA", "DMUL", "DADD", "DSETP", "DMNMX"}
def main(path, sub, top=45):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, cur, seen = None, None, 0
    by_op = collections.Counter(); thr_op = collections.Counter(); by_line = collections.defaultdict(lambda: [0, 0, 0, 0])
    for r in rows:
        if len(r) >= 2 and r[0] in ("Function Name", "Kernel Name"):
            cur = r[1]; seen += sub in cur
            if seen > 1 and sub in cur: cur = None
            continue
        if r and r[0] == "Address": hdr = r; continue
        if hdr and cur and sub in cur and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            src = d.get("Source", "").strip().split()
            if not src: continue
            op = src[1] if src[0].startswith("@") and len(src) > 1 else src[0]
            op = op.split(".")[0]
            try:
                n = int(d.get("Instructions Executed") or 0); t = int(d.get("Predicated-On Thread Instructions Executed") or 0)
                s = int(d.get("# Samples") or 0)
            except ValueError: continue
            by_op[op] += n; thr_op[op] += t
            loc = d.get("Source Location") or d.get("Location") or "?"
            e = by_line[loc]; e[0] += n; e[1] += n if op in FP64 else 0; e[2] += s; e[3] += t
    tot = sum(by_op.values()); f64 = sum(v for k, v in by_op.items() if k in FP64)
    print(f"warp instructions {tot}, FP64-pipe {f64} ({100*f64/tot:.1f}%), thread instr {sum(thr_op.values())}")
    for op, n in by_op.most_common(28): print(f"  {op:10s} {n:12d} {100*n/tot:5.1f}%  lanes {thr_op[op]/max(n,1):5.1f}")
    print("by source line (warp instr %, fp64 share, samples %):")
    ts = sum(e[2] for e in by_line.values())
    for loc, e in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {loc[-40:]:>40s} inst {100*e[0]/tot:5.1f}% fp64 {100*e[1]/max(e[0],1):4.0f}% smp {100*e[2]/max(ts,1):5.1f}% lanes {e[3]/max(e[0],1):5.1f}")
if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 45)
