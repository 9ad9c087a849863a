#!/bin/bash
# A/B of library builds on the device-resident step (developer tool): bash tools/ab_kernels.sh <lib.so> [<lib.so> ...]
# Each build runs tests/tools/quick_time.py (CUDA events per kernel stage, flags checked against the oracle on a sample).
for lib in "$@"; do
  for wl in ${WLS:-cfg2 cfg3 cfg4 cfg5}; do
    mode=epa; [ $wl = cfg2 ] && mode=gjk
    echo "== $lib $wl"
    GJKEPA_LIB=$lib WL=$wl MODE=$mode LANES=0 CHECK=${CHECK:-20000} python tests/tools/quick_time.py 2>&1 | grep quick
  done
done
