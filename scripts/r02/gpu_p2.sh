#!/bin/bash
# Round 2, call P2: the general pushers through whole intervals of gpat_particle_mover (cell-sorted particles), both routings
mkdir -p gpurun_out
T=r02p2
ALT_PROBE_MODE=mover ALT_PROBE_ROUTES=1,0 timeout 1500 python scripts/r02/alt_probe.py 2>&1 | tee gpurun_out/${T}_alt_probe_mover.log
