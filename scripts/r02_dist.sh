set -x
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_dist_nccl.py -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -5 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print({k:d.get(k) for k in ('value','ms_per_step','build_ms','apply_fwd_ms','apply_T_ms','gpu_launches','parity')})
print(d.get('collective_ms')); print(d['e2e'])
print(d.get('build_phases_ms'))
PY
