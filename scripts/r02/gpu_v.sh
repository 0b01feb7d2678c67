#!/bin/bash
# Round 2, call V: kSpecAlt instantiations (1-D, focused transport, maps behind the lane-group gather): GPU suite + throughput A/B
mkdir -p gpurun_out
T=r02v
python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -5 gpurun_out/${T}_pytest_gpu.log
timeout 600 python scripts/r02/alt_probe.py 200 > gpurun_out/${T}_alt_probe.log 2>&1; cat gpurun_out/${T}_alt_probe.log
