# final build: pair-kernel ncu capture, bench lines, GPU tests
ncu --set full --clock-control none --import-source on -s 2 -c 1 -f -k regex:blur_adjsig_pair -o gpurun_out/prof_r2b_pair python profiles/run_kernels.py --only blur_pair --batch 8 > gpurun_out/ncu_r2b_b.log 2>&1
python bench.py --steps 200 --warmup 5 > gpurun_out/bench_r2b_n1_f16.json 2> gpurun_out/bench_r2b_n1_f16.err
python bench.py --workload f4 --steps 50 --warmup 5 > gpurun_out/bench_r2b_n1_f4.json 2> gpurun_out/bench_r2b_n1_f4.err
python bench.py --workload celeba --steps 50 --warmup 5 > gpurun_out/bench_r2b_n1_celeba.json 2> gpurun_out/bench_r2b_n1_celeba.err
python -m pytest tests -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_r2b_gpu.txt
tail -2 gpurun_out/pytest_r2b_gpu.txt
