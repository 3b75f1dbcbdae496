# ncu --set full captures of the three level-0 kernels (1024 maps of 256^2) and the launch list of eager steps
NCU="ncu --set full --clock-control none --import-source on -s 2 -c 1 -f"
$NCU -k regex:ffl_kernel -o gpurun_out/prof_r2b_ffldiff python profiles/run_kernels.py --only ffl_diff --batch 8 > gpurun_out/ncu_r2b_a.log 2>&1
$NCU -k regex:blur_adjsig_pair -o gpurun_out/prof_r2b_pair python profiles/run_kernels.py --only blur_pair --batch 8 > gpurun_out/ncu_r2b_b.log 2>&1
$NCU -k regex:blur_diff -o gpurun_out/prof_r2b_diff python profiles/run_kernels.py --only blur_diff --batch 8 > gpurun_out/ncu_r2b_c.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 2 --warmup 3 --no-graph > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out/prof_r2b_* gpurun_out/launches_r2b.csv
