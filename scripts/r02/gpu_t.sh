#!/bin/bash
# Round 2, call T: all conversions on the integer pipe (GPAT_CVT_ALU_MASK=15) vs the default (10).
mkdir -p gpurun_out
T=r02t
V=$PWD/stochastic_parker_b200/csrc/libgpat_cuda.cvt15.so
GPAT_LIB=$V timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-membw --no-strong > gpurun_out/${T}_c1_mask15.json 2> gpurun_out/${T}_c1_mask15.err
python -c "
import json;d=json.load(open('gpurun_out/${T}_c1_mask15.json'));print('c1 mask15 value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))"
python scripts/r02/c5_probe.py 256 16000000 "mask10:" > gpurun_out/${T}_c5_mask10.log 2>&1; tail -1 gpurun_out/${T}_c5_mask10.log
GPAT_LIB=$V python scripts/r02/c5_probe.py 256 16000000 "mask15:" > gpurun_out/${T}_c5_mask15.log 2>&1; tail -1 gpurun_out/${T}_c5_mask15.log
GPAT_LIB=$V python scripts/r02/c5_probe.py 512 125000000 "mask15:" > gpurun_out/${T}_c5_512_mask15.log 2>&1; tail -1 gpurun_out/${T}_c5_512_mask15.log
