# L2E (D_pp) kernel at 3 (default) vs 4 resident CTAs per SM (make variant TAG=mb4 EXTRA=-DGPAT_MINBLOCKS=4)
mkdir -p gpurun_out
{
echo "== default"; timeout 60 python scripts/c4_probe.py c4 1024 400000 2
echo "== mb4"; GPAT_LIB=$PWD/stochastic_parker_b200/csrc/libgpat_cuda.mb4.so timeout 60 python scripts/c4_probe.py c4 1024 400000 2
} > gpurun_out/c4_mb4.log 2>&1
cat gpurun_out/c4_mb4.log
