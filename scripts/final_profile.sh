# Final measurement + profile pass of a round (run under gpurun; writes into gpurun_out/).
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
for w in cfg4 cfg3 cfg2 cfg1; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; done
python scripts/bench_configs.py > gpurun_out/configs.txt 2>&1
python scripts/bench_spmm.py > gpurun_out/spmm.txt 2>&1
python scripts/bench_apply.py > gpurun_out/apply.txt 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"clip_quad_kernel" -s 1 -c 1 -o gpurun_out/clip_final -f python scripts/prof_build.py > gpurun_out/ncu_clip.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"spmv_sell" -c 2 -o gpurun_out/spmv_final -f python scripts/prof_build.py > gpurun_out/ncu_spmv.log 2>&1
timeout 500 ncu --set full --clock-control none -k regex:"rs_downsweep_kernel|row_sort_split_kernel|sell_fill_kernel|bp_bounds_kernel|bp_bin_kernel|bp_query_kernel" -s 11 -c 11 -o gpurun_out/build_final -f python scripts/prof_build.py > gpurun_out/ncu_asm.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -n 3 gpurun_out/configs.txt; tail -n 3 gpurun_out/spmm.txt; tail -n 3 gpurun_out/apply.txt
