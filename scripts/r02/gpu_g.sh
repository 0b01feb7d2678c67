#!/bin/bash
# Round 2, call G: tiled sort keys and the 3-D switch-specialised kernel on C5 (256^3 and full size), one process per grid.
mkdir -p gpurun_out
T=r02g
python -m pytest tests -m gpu -q -x -k "side_plane or golden or sorted or c5" > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -3 gpurun_out/${T}_pytest_gpu.log
python scripts/r02/c5_probe.py 256 16000000 "tile64x32x32+spec:" "tile+nospec:GPAT_NO_SPEC3D=1" "untiled+spec:GPAT_SORT_TILE=0" "tile32^3+spec:GPAT_SORT_TILE=32,32,32" > gpurun_out/${T}_c5_probe_256.log 2>&1
cat gpurun_out/${T}_c5_probe_256.log
python scripts/r02/c5_probe.py 512 125000000 "tile64x32x32+spec:" "untiled+spec:GPAT_SORT_TILE=0" "tile32^3+spec:GPAT_SORT_TILE=32,32,32" "tile128x64x32:GPAT_SORT_TILE=128,64,32" "tile+nospec:GPAT_NO_SPEC3D=1" "tile+L3B:GPAT_NO_L3D=1" > gpurun_out/${T}_c5_probe_512.log 2>&1
cat gpurun_out/${T}_c5_probe_512.log
