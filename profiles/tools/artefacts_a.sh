python bench.py --steps 200 --warmup 5 > gpurun_out/bench_r2b_n1_f16.json 2> gpurun_out/bench_r2b_n1_f16.err
python bench.py --workload f4 --steps 50 --warmup 5 > gpurun_out/bench_r2b_n1_f4.json 2> gpurun_out/bench_r2b_n1_f4.err
python bench.py --workload celeba --steps 50 --warmup 5 > gpurun_out/bench_r2b_n1_celeba.json 2> gpurun_out/bench_r2b_n1_celeba.err
python -m pytest tests -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_r2b_gpu.txt
( time timeout 600 python bench.py --microbench > gpurun_out/microbench_r2b.txt 2> gpurun_out/microbench_r2b.err ) 2> gpurun_out/microbench_time.txt
tail -3 gpurun_out/pytest_r2b_gpu.txt; tail -3 gpurun_out/microbench_time.txt
