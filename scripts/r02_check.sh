# Round-2 check pass: GPU tests, build phases, compute-sanitizer on the smoke script (run under gpurun).
set -x
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for w in cfg5 cfg2 cfg1; do timeout 300 python scripts/time_build.py $w 12; done
for tool in memcheck racecheck initcheck; do timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_smoke.py > gpurun_out/sanitize_$tool.log 2>&1; tail -3 gpurun_out/sanitize_$tool.log; done
