mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "tracking or restart or step_parity or interval_parity" ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -1
for t in "" m0 m2 m15; do
lib=stochastic_parker_b200/csrc/libgpat_cuda${t:+.$t}.so
GPAT_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/bench6_${t:-def}.json 2> gpurun_out/bench6_${t:-def}.err
python -c "
import json;d=json.load(open('gpurun_out/bench6_${t:-def}.json'));print('c1 ${t:-def} value %.4g push_ms %.2f' % (d['value'], d['breakdown_ms_per_step']['push_ms']))"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:push_kernel -s 1 -c 1 -o gpurun_out/prof_c1f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c1f.log 2>&1
tail -1 gpurun_out/ncu_c1f.log | cut -c1-100
