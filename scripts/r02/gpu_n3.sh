#!/bin/bash
# Round 2, call N3: ncu --set full of the final general-pusher kernels in a whole cell-sorted interval (second mover call):
# focused transport 2-D and 2-D Parker + maps
mkdir -p gpurun_out
T=r02n3
for k in ft_2d_dpp maps_2d; do
  ALT_PROBE_MODE=mover ALT_PROBE_ONLY=$k ALT_PROBE_ROUTES=0 ncu --set full --clock-control none -k regex:push_kernel_coop -s 1 -c 1 -f -o /tmp/${T}_$k python scripts/r02/alt_probe.py > gpurun_out/${T}_ncu_$k.log 2>&1
  ncu -i /tmp/${T}_$k.ncu-rep --page raw --csv > /tmp/${T}_${k}_raw.csv 2>/dev/null
  python scripts/ncu_keys.py /tmp/${T}_${k}_raw.csv > gpurun_out/${T}_push_coop_${k}_ncu.txt 2>&1
  head -4 gpurun_out/${T}_push_coop_${k}_ncu.txt
done
