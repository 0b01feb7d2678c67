# C4 probe (see scripts/c4_probe.py); results -> gpurun_out/c4_probe.log
mkdir -p gpurun_out
{
timeout 100 python scripts/c4_probe.py c4 1024 400000 2
timeout 150 python scripts/c4_probe.py c4 4096 170000 2
} > gpurun_out/c4_probe.log 2>&1
cat gpurun_out/c4_probe.log
