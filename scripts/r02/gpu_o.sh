#!/bin/bash
# Round 2, call O: rolled round loop in the 3-D kernels (instruction-fetch stalls) A/B against the unrolled build.
mkdir -p gpurun_out
T=r02o
python -m pytest tests -m gpu -q -x -k "side_plane or golden or step_parity or interval_parity or c5 or 3d" > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -3 gpurun_out/${T}_pytest_gpu.log
U=$PWD/stochastic_parker_b200/csrc/libgpat_cuda.unroll.so
{
echo "== rolled (default), 256^3"; python scripts/r02/c5_probe.py 256 16000000 "rolled:"
echo "== unrolled, 256^3"; GPAT_LIB=$U python scripts/r02/c5_probe.py 256 16000000 "unrolled:"
echo "== rolled (default), 512^3"; python scripts/r02/c5_probe.py 512 125000000 "rolled:" "rolled_maxctas3:GPAT_PUSH_MAXCTAS=3"
echo "== unrolled, 512^3"; GPAT_LIB=$U python scripts/r02/c5_probe.py 512 125000000 "unrolled:"
} > gpurun_out/${T}_c5_rolled.log 2>&1
grep -E "^==|steps/s" gpurun_out/${T}_c5_rolled.log
