#!/bin/bash
# Round 2, call M: corrected gather micro-benchmark (index by mask, eight loads in flight), all four patterns.
mkdir -p gpurun_out
python - > gpurun_out/r02m_membw.log 2>&1 <<'PY'
import sys
sys.path.insert(0, ".")
import bench, json
print(json.dumps(bench.live_membw(True), indent=1))
PY
cat gpurun_out/r02m_membw.log
