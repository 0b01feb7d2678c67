#!/bin/bash
# Round 2, call Q2: one kSpecAlt instantiation per pusher -- parity tests of the general pushers + both probes
mkdir -p gpurun_out
T=r02q2
python -m pytest tests -m gpu -q -k "general or maps or step_parity or tracking or golden or histograms" > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -4 gpurun_out/${T}_pytest_gpu.log
ALT_PROBE_ROUTES=0 python scripts/r02/alt_probe.py 200 2>&1 | tee gpurun_out/${T}_alt_probe.log
ALT_PROBE_MODE=mover ALT_PROBE_ROUTES=0 timeout 1500 python scripts/r02/alt_probe.py 2>&1 | tee gpurun_out/${T}_alt_probe_mover.log
