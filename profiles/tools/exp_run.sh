#!/bin/bash
# usage: exp_run.sh <outfile> <variant>:<only> ...   (variant "stock" = the in-tree library)
out=gpurun_out/$1; shift
mkdir -p gpurun_out; : > $out
for spec in "$@"; do
  v=${spec%%:*}; only=${spec##*:}
  echo "== $v ($only)" >> $out
  if [ "$v" = stock ]; then
    timeout 300 python profiles/run_kernels.py --only $only --batch 32 --iters 20 >> $out 2>&1
  else
    FAVAE_B200_LIB=$PWD/profiles/tools/variants/lib_$v.so timeout 300 python profiles/run_kernels.py --only $only --batch 32 --iters 20 >> $out 2>&1
  fi
done
cat $out
