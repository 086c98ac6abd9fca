"""Build / time variants of the library that differ by -D flags (kernel tuning experiments).
  python scripts/variants.py build NAME=-DFLAG=1,-DOTHER=2 ...     (here, no GPU needed)
  python scripts/variants.py run [workload]                         (on the GPU box: times every built variant)"""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "conservativeregridding.jl_b200", "csrc", "variants")
def build(specs):
    os.makedirs(VDIR, exist_ok=True)
    procs = []
    for spec in specs:
        name, _, flags = spec.partition("=")
        env = dict(os.environ, CRG_LIB=os.path.join(VDIR, f"libcrgb200_{name}.so"), CRG_NVCC_EXTRA=flags.replace(",", " "))
        procs.append((name, subprocess.Popen([sys.executable, "-c", "import sys; sys.path.insert(0, %r); from crg_b200 import _lib; _lib.build(force=True)" % ROOT], env=env)))
    for name, p in procs:
        print(name, "rc", p.wait())
def run(workload):
    for lib in ["mode0", None] + sorted(glob.glob(os.path.join(VDIR, "libcrgb200_*.so"))):
        env = dict(os.environ)
        if lib == "mode0": env["CRG_CLIP_MODE"] = "0"; lib = None
        if lib: env["CRG_LIB"] = lib
        out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "time_build.py"), workload], env=env, capture_output=True, text=True)
        print(os.path.basename(lib) if lib else "default", out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:], flush=True)
if __name__ == "__main__":
    if sys.argv[1] == "build": build(sys.argv[2:])
    else: run(sys.argv[2] if len(sys.argv) > 2 else "cfg5")
