set -x
for N in ${NS:-8 4 2}; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 $BENCH_EXTRA > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -2 gpurun_out/bench_n$N.err
done
python - <<'PY'
import json
for N in (2,4,8):
    try:
        d=json.loads([l for l in open(f'gpurun_out/bench_n{N}.json') if l.startswith('{')][-1])
        print(N, {k:d.get(k) for k in ('value','ms_per_step','build_ms','apply_fwd_ms','apply_T_ms')}, d.get('build_phases_ms'), d.get('collective_ms'), d['e2e']['ms_per_step'], d.get('per_rank'))
    except Exception as e: print(N, 'ERR', e)
PY
