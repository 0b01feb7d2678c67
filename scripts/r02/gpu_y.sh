#!/bin/bash
# Round 2, call Y: lane-group gather of the turbulence-map record (kSpecAltMaps instantiations): GPU suite + throughput probe
mkdir -p gpurun_out
T=${T:-r02y4}
python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -5 gpurun_out/${T}_pytest_gpu.log
ALT_PROBE_ROUTES=0 python scripts/r02/alt_probe.py 200 > gpurun_out/${T}_alt_probe.log 2>&1; cat gpurun_out/${T}_alt_probe.log
