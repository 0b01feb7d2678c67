#!/bin/bash
# Round 2, call T3: skipping rounds whose owner lanes are idle (the tail of a launch), A/B against the same build without
mkdir -p gpurun_out
T=r02t3
CS=stochastic_parker_b200/csrc
for v in ${VARIANTS:-default noskip default noskip}; do
  if [ $v = default ]; then unset GPAT_LIB; else export GPAT_LIB=$PWD/$CS/libgpat_cuda.$v.so; fi
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-membw --no-strong > gpurun_out/${T}_c1_$v.json 2>/dev/null
  python bench.py --nptl 125000 --steps 6 --warmup 3 --no-cpu-baseline --no-membw --no-strong > gpurun_out/${T}_c1s_$v.json 2>/dev/null
  python bench.py --workload c4 --grid 1024 --nptl 2000000 --steps 1 --warmup 1 --no-cpu-baseline --no-membw > gpurun_out/${T}_c4_$v.json 2>/dev/null
  python bench.py --workload c3 --steps 4 --warmup 3 --no-cpu-baseline --no-membw > gpurun_out/${T}_c3_$v.json 2>/dev/null
  python - <<PY
import json
o = []
for k in ("c1", "c1s", "c4", "c3"):
    try: o.append("%s %.4g" % (k, json.load(open("gpurun_out/${T}_%s_$v.json" % k))["value"]))
    except Exception as e: o.append("%s failed" % k)
print("$v", *o, flush=True)
PY
  python scripts/r02/c5_probe.py 256 16000000 "$v:" 2>&1 | tail -1
done 2>&1 | tee gpurun_out/${T}_summary.log
