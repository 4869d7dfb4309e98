#!/bin/bash
# A/B of library variants on config 4 (65536 chains): thread-per-chain persistent kernel
for lib in aehmc_b200/lib "$@"; do
  echo "== $lib"
  B2H_LIB=$lib/libb200hmc.so python benchmarks/c4_tail_probe.py 2>&1 | grep "G=1"
done
