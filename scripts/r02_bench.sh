set -x
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','build_ms','build_ms_median','build_ms_best','apply_fwd_ms','apply_T_ms','gpu_launches')})
print(d['build_phases_ms']); print(d['e2e']['ms_per_step'], d['e2e']['host_ms']); print(d['e2e_explicit_cells']['ms_per_step'], d['e2e_explicit_cells']['host_ms'])
PY
