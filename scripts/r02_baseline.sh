# Round-2 baseline pass (run under gpurun): GPU tests, bench line, per-config build phases, launch list, full ncu of the clip kernel.
set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
for w in cfg5 cfg2 cfg1 cfg3 cfg4; do timeout 300 python scripts/time_build.py $w 12; done > gpurun_out/time_build.txt 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"clip_quad_kernel" -s 1 -c 1 -o gpurun_out/clip_base -f python scripts/prof_build.py > gpurun_out/ncu_clip.log 2>&1
cat gpurun_out/time_build.txt
