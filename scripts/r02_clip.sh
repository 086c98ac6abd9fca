# Round-2 clip pass: GPU tests, per-config build phases, ncu of the clip kernels (run under gpurun).
set -x
python scripts/debug_clip.py 2>&1 | tail -12
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for w in cfg5 cfg2 cfg1 cfg3 cfg4; do timeout 300 python scripts/time_build.py $w 12; done > gpurun_out/time_build.txt 2>&1
cat gpurun_out/time_build.txt
python scripts/variants.py run cfg5
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"clip_quad|quad_normals" -s 2 -c 2 -o gpurun_out/clip_fast -f python scripts/prof_build.py > gpurun_out/ncu_clip.log 2>&1
tail -3 gpurun_out/ncu_clip.log
