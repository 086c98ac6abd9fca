# Final measurement + profile pass of a round (run under gpurun; writes into gpurun_out/).
set -x
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python scripts/bench_configs.py > gpurun_out/configs.txt 2>&1
python scripts/bench_spmm.py > gpurun_out/spmm.txt 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"clip_quad_kernel" -s 1 -c 1 -o gpurun_out/clip_final -f python scripts/prof_build.py > gpurun_out/ncu_clip.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"spmv_sell" -c 2 -o gpurun_out/spmv_final -f python scripts/prof_build.py > gpurun_out/ncu_spmv.log 2>&1
timeout 500 ncu --set full --clock-control none -k regex:"rs_downsweep_kernel|row_sort_split_kernel|sell_fill_kernel" -s 3 -c 5 -o gpurun_out/asm_final -f python scripts/prof_build.py > gpurun_out/ncu_asm.log 2>&1
ls -la gpurun_out/*.ncu-rep
