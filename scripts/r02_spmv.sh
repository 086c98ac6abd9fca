set -x
for lib in "" conservativeregridding.jl_b200/csrc/variants/libcrgb200_nt128.so conservativeregridding.jl_b200/csrc/variants/libcrgb200_mb4.so; do
for deep in 1 2; do for bps in 6 8 10 16 24 48; do echo "== lib=$lib deep=$deep bps=$bps"; CRG_LIB=$lib CRG_SELL_DEEP=$deep CRG_SELL_BPS=$bps timeout 300 python scripts/bench_apply.py 2>&1 | grep -E "T=True"; done; done; done
