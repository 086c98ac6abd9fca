set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 --workload cfg5x4 > gpurun_out/bench_big_n8.json 2> gpurun_out/bench_big_n8.err; tail -2 gpurun_out/bench_big_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 10 --warmup 3 --workload cfg5x4 > gpurun_out/bench_big_n4.json 2> gpurun_out/bench_big_n4.err; tail -2 gpurun_out/bench_big_n4.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload cfg5x4 --no-cpu-baseline > gpurun_out/bench_big_n1.json 2> gpurun_out/bench_big_n1.err; tail -2 gpurun_out/bench_big_n1.err
python - <<'PY'
import json
for N in (1,4,8):
    try:
        d=json.loads([l for l in open(f'gpurun_out/bench_big_n{N}.json') if l.startswith('{')][-1])
        print(N, {k:d.get(k) for k in ('value','ms_per_step','build_ms','apply_fwd_ms','apply_T_ms')}, d.get('build_phases_ms'), d.get('collective_ms'), d['e2e']['ms_per_step'], d.get('per_rank'))
    except Exception as e: print(N, 'ERR', e)
PY
