#!/bin/bash
# A/B of the FP64 dense-apply kernel builds (benchmarks/dense_apply_ab.py at the c2 shape):
#   head = previous commit's kernel, f0 = 64-bit fragment loads + lean copy issue, main = 128-bit fragment loads,
#   ns = 128-bit fragment loads, every warp refills right after the barrier
mkdir -p gpurun_out
for v in head f0 ns main; do
  if [ "$v" = main ]; then unset B2H_LIB; else export B2H_LIB=$PWD/build/lib_$v/libb200hmc.so; fi
  [ "$v" != main ] && [ ! -f "$B2H_LIB" ] && continue
  echo "== $v"
  python benchmarks/dense_apply_ab.py 2>&1 | tail -3
done
