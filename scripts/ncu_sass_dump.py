"""Dump the SASS of one kernel of an .ncu-rep with executed warp-instruction counts, active lanes and samples per instruction.
usage: ncu_sass_dump.py rep kernel_substr > out.txt"""
import csv, io, subprocess, sys
def main(path, sub):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, cur, seen = None, None, 0
    for r in rows:
        if len(r) >= 2 and r[0] in ("Function Name", "Kernel Name"):
            cur = r[1]; seen += sub in cur
            if seen > 1 and sub in cur: cur = None
            continue
        if r and r[0] == "Address": hdr = r; continue
        if hdr and cur and sub in cur and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            n = int(d.get("Instructions Executed") or 0); t = int(d.get("Predicated-On Thread Instructions Executed") or 0)
            s = int(d.get("# Samples") or 0)
            print(f"{d['Address'][-5:]} {n:10d} {t/max(n,1):5.1f} {s:6d}  {d['Source'].strip()}")
if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
