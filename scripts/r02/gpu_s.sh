#!/bin/bash
# Round 2, call S2: compute-sanitizer (memcheck, racecheck) over the kSpecAlt / kSpecAltMaps kernels through their parity tests
mkdir -p gpurun_out
T=r02s2
K='test_turbulence_maps_step_parity or (test_step_parity and 0- and (focused or ft or shock_1d))'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -k "$K" > gpurun_out/${T}_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/${T}_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -k "$K" > gpurun_out/${T}_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/${T}_sanitizer_racecheck.log
tail -4 gpurun_out/${T}_sanitizer_memcheck.log gpurun_out/${T}_sanitizer_racecheck.log
