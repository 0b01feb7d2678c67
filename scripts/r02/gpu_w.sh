#!/bin/bash
# Round 2, call W: ncu --set full with source counters of (1) the C1 production kernel, (2) the focused-transport kSpecAlt kernel
# (the .ncu-rep files are 33 MB each: dumped to CSV on the box, only the CSVs come back)
mkdir -p gpurun_out
T=r02w
ncu --set full --clock-control none --import-source on -k regex:push_kernel_coop -s 2 -c 1 -f -o /tmp/${T}_c1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-membw --no-strong > gpurun_out/${T}_ncu_c1.log 2>&1
ALT_PROBE_ONLY=ft_2d_dpp GPAT_ALT_STRICT=0 ncu --set full --clock-control none --import-source on -k regex:push_kernel_coop -s 1 -c 1 -f -o /tmp/${T}_ft python scripts/r02/alt_probe.py 100 > gpurun_out/${T}_ncu_ft.log 2>&1
for k in c1 ft; do
  ncu -i /tmp/${T}_$k.ncu-rep --page raw --csv > gpurun_out/${T}_${k}_raw.csv 2>/dev/null
  ncu -i /tmp/${T}_$k.ncu-rep --page source --csv > gpurun_out/${T}_${k}_source.csv 2>/dev/null
done
ls -la gpurun_out/${T}_*
