#!/bin/bash
# Round 2, call S: all FP32->FP64 conversions on the XU (GPAT_CVT_ALU_MASK=0) vs half on the integer pipe (default, 10).
mkdir -p gpurun_out
T=r02s
run() { n=$1; e=$2; shift 2
  env $e timeout 900 python bench.py "$@" --no-cpu-baseline --no-membw --no-strong > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err
  python -c "
import json;d=json.load(open('gpurun_out/${T}_$n.json'));print('$n value %.4g e2e %.4g push_ms %.2f clocks %s' % (d['value'], d['e2e']['value'], d['breakdown_ms_per_step']['push_ms'], d['clocks'].get('sm_mhz')))" || tail -3 gpurun_out/${T}_$n.err
}
V=GPAT_LIB=$PWD/stochastic_parker_b200/csrc/libgpat_cuda.cvt0.so
run c1_mask10 "X=1" --steps 6 --warmup 3
run c1_mask0 "$V" --steps 6 --warmup 3
run c4_mask10 "X=1" --workload c4 --grid 1024 --nptl 2000000 --steps 1 --warmup 1
run c4_mask0 "$V" --workload c4 --grid 1024 --nptl 2000000 --steps 1 --warmup 1
python scripts/r02/c5_probe.py 256 16000000 "mask10:" > gpurun_out/${T}_c5_mask10.log 2>&1; tail -1 gpurun_out/${T}_c5_mask10.log
GPAT_LIB=$PWD/stochastic_parker_b200/csrc/libgpat_cuda.cvt0.so python scripts/r02/c5_probe.py 256 16000000 "mask0:" > gpurun_out/${T}_c5_mask0.log 2>&1; tail -1 gpurun_out/${T}_c5_mask0.log
python scripts/r02/c5_probe.py 512 125000000 "mask10:" > gpurun_out/${T}_c5_512_mask10.log 2>&1; tail -1 gpurun_out/${T}_c5_512_mask10.log
GPAT_LIB=$PWD/stochastic_parker_b200/csrc/libgpat_cuda.cvt0.so python scripts/r02/c5_probe.py 512 125000000 "mask0:" > gpurun_out/${T}_c5_512_mask0.log 2>&1; tail -1 gpurun_out/${T}_c5_512_mask0.log
