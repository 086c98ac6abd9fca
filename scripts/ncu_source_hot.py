"""Per-source-line hot spots of an .ncu-rep captured with --import-source on (needs -lineinfo):
samples, instructions, average active threads, dominant stall reasons.  usage: ncu_source_hot.py rep [topN]"""
import csv, io, subprocess, sys
def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, hdr, recs = None, None, []
    for r in rows:
        if len(r) == 2 and r[0] in ("File Path", "File Name"): cur_file = r[1].split("/")[-1]; continue
        if len(r) == 2: continue
        if r and r[0] == "Line No": hdr = r; continue
        if hdr and len(r) == len(hdr) and r[0].isdigit():
            d = dict(zip(hdr, r)); d["file"] = cur_file; recs.append(d)
    tot_s = sum(int(d["# Samples"] or 0) for d in recs); tot_i = sum(int(d["Instructions Executed"] or 0) for d in recs)
    print(f"total samples {tot_s}, warp instructions {tot_i}")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    recs.sort(key=lambda d: -int(d["# Samples"] or 0))
    for d in recs[:top]:
        s = int(d["# Samples"] or 0)
        st = sorted(((int(d[h] or 0), h[6:]) for h in stalls), reverse=True)[:3]
        print(f"{d['file']}:{d['Line No']:>4} smp {100*s/tot_s:5.1f}% inst {100*int(d['Instructions Executed'] or 0)/tot_i:5.1f}% thr {float(d['Avg. Threads Executed'] or 0)/max(1,int(d['Instructions Executed'] or 1))*int(d['Instructions Executed'] or 1) if False else d['Avg. Threads Executed']:>6} "
              + " ".join(f"{n}:{100*v/max(s,1):.0f}%" for v, n in st) + "  | " + d["Source"].strip()[:110])
if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
