# one GPU call: parity tests, headline bench, the other named configs, ncu of the HBM-resident configs
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
cut -c1-400 gpurun_out/bench_c1.json
bash scripts/gpu_configs.sh
for spec in "c4 --nptl 200000" "c5 --grid 256 --nptl 200000"; do
  set -- $spec; wl=$1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:push_kernel -s 1 -c 1 -o gpurun_out/prof_$wl python bench.py --workload $spec --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$wl.log 2>&1
  tail -1 gpurun_out/ncu_$wl.log | cut -c1-200
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
