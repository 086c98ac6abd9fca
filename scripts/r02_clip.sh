# Round-2 clip pass (opt-in wedge path): GPU tests, per-config build phases, kernel durations (run under gpurun).
set -x
export CRG_CLIP_FAST=1
python scripts/debug_clip.py 2>&1 | tail -6
timeout 1500 python -m pytest tests/test_gpu_build.py tests/test_gpu_parity_full.py tests/test_gpu_highprec.py tests/test_gpu_sweat.py -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for w in cfg5 cfg2 cfg1 cfg3 cfg4; do timeout 300 python scripts/time_build.py $w 12; done > gpurun_out/time_build.txt 2>&1
cat gpurun_out/time_build.txt
python scripts/variants.py run cfg5
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"clip_|quad_normals" -s 4 -c 4 --csv --log-file gpurun_out/clip_launches.csv python scripts/prof_build.py > /dev/null 2>&1
cut -d, -f5,12- gpurun_out/clip_launches.csv | tail -6
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"clip_|quad_normals" -s 4 -c 4 -o gpurun_out/clip_split -f python scripts/prof_build.py > gpurun_out/ncu_clip.log 2>&1
tail -2 gpurun_out/ncu_clip.log
