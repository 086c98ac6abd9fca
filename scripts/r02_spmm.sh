set -x
timeout 600 python -m pytest tests/test_gpu_apply.py -q 2>&1 | tail -2
for pair in cfg3 healpix lonlat; do for t in 1 0; do echo "== $pair tiled=$t"; PAIR=$pair CRG_SPMM_TILED=$t python scripts/bench_spmm.py | grep -E "cell-fastest"; done; done
