#!/bin/bash
# Round 2, call K2: log / exp constants from a __constant__ table in the general-pusher kernels (381 UMOV in the FT kernel)
mkdir -p gpurun_out
T=r02k2
CS=stochastic_parker_b200/csrc
for v in default ktab default ktab; do
  if [ $v = default ]; then unset GPAT_LIB; else export GPAT_LIB=$PWD/$CS/libgpat_cuda.$v.so; fi
  echo "== $v"
  ALT_PROBE_ROUTES=0 python scripts/r02/alt_probe.py 200 2>&1 | grep -v plain
done 2>&1 | tee gpurun_out/${T}_alt_probe.log
