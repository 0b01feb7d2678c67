#!/bin/bash
# Round 2, call N: ncu --set full of the push kernel on C5 at its stated size (512^3, 1.25e8 particles), second interval.
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none -k regex:push_kernel -s 1 -c 1 -o gpurun_out/r02n_prof_c5_512 python scripts/r02/c5_probe.py 512 125000000 "default:" > gpurun_out/r02n_ncu_c5_512.log 2>&1
tail -3 gpurun_out/r02n_ncu_c5_512.log
