#!/bin/bash
# Round 2, call N2: ncu --set full of the general-pusher kernels in the production build (third launch of the probe, 400 pushes
# per particle): focused transport 2-D, 1-D, 2-D + maps, focused transport 3-D.  Raw pages dumped on the box.
mkdir -p gpurun_out
T=r02n2
for k in ft_2d_dpp shock_1d maps_2d ft_3d; do
  ALT_PROBE_ONLY=$k ALT_PROBE_ROUTES=0 ncu --set full --clock-control none -k regex:push_kernel_coop -s 2 -c 1 -f -o /tmp/${T}_$k python scripts/r02/alt_probe.py 400 > gpurun_out/${T}_ncu_$k.log 2>&1
  ncu -i /tmp/${T}_$k.ncu-rep --page raw --csv > gpurun_out/${T}_${k}_raw.csv 2>/dev/null
  python scripts/ncu_keys.py gpurun_out/${T}_${k}_raw.csv > gpurun_out/${T}_push_coop_${k}_ncu.txt 2>&1
  head -3 gpurun_out/${T}_push_coop_${k}_ncu.txt
done
