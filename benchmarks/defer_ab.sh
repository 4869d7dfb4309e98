#!/bin/bash
# A/B of the deferred rare paths in the thread-per-chain persistent kernel (c4, 65536 chains)
for lib in aehmc_b200/lib build/lib_nodefer build/lib_l24t3 build/lib_l32t4 build/lib_l16t1; do
  echo "== $lib"
  B2H_LIB=$lib/libb200hmc.so python benchmarks/c4_tail_probe.py 2>&1 | grep "G="
done
