set -x
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for w in cfg5 cfg2 cfg1 cfg3 cfg4; do timeout 300 python scripts/time_build.py $w 12; done
