#!/bin/bash
# pair-kernel rows-per-iteration experiment: one microbench line per variant library
mkdir -p gpurun_out
out=gpurun_out/exp_pair_u.txt
: > $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $out
for u in 0 1 2 3 4 0; do
  echo "== FAVAE_PAIR_U=$u" >> $out
  FAVAE_B200_LIB=$PWD/profiles/tools/variants/lib_pairU$u.so timeout 300 python profiles/run_kernels.py --only blur_pair --batch 32 --iters 20 >> $out 2>&1
done
echo "== ffl default" >> $out
timeout 300 python profiles/run_kernels.py --only ffl --batch 32 --iters 10 >> $out 2>&1
echo "== ffl c4" >> $out
FAVAE_FFL256=c4 timeout 300 python profiles/run_kernels.py --only ffl --batch 32 --iters 10 >> $out 2>&1
cat $out
