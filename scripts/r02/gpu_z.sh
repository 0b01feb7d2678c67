#!/bin/bash
# Round 2, call Z: resident CTAs per SM of the kSpecAlt instantiations (2 / 3 / 4 = 255 / 168 / 128 registers)
mkdir -p gpurun_out
T=r02z
CS=stochastic_parker_b200/csrc
for v in default amb2 amb4; do
  if [ $v = default ]; then unset GPAT_LIB; else export GPAT_LIB=$PWD/$CS/libgpat_cuda.$v.so; fi
  echo "== $v" | tee -a gpurun_out/${T}_alt_probe.log
  ALT_PROBE_ROUTES=0 python scripts/r02/alt_probe.py 200 2>&1 | tee -a gpurun_out/${T}_alt_probe.log
done
