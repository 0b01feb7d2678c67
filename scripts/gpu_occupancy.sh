# occupancy sweep of the push kernel: pad dynamic smem so that 1..4 CTAs fit per SM
mkdir -p gpurun_out
for pad in 200000 90000 50000 0; do
  GPAT_PUSH_SMEM_PAD=$pad python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_pad$pad.json 2>> gpurun_out/occ.log
  python -c "
import json;d=json.load(open('gpurun_out/bench_pad$pad.json'));print('pad $pad', '%.4g' % d['value'], 'push_ms %.2f' % d['breakdown_ms_per_step']['push_ms'])"
done
for t in "$@"; do
  GPAT_LIB=$PWD/stochastic_parker_b200/csrc/libgpat_cuda.$t.so python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_$t.json 2>> gpurun_out/occ.log
  python -c "
import json;d=json.load(open('gpurun_out/bench_$t.json'));print('$t', '%.4g' % d['value'], 'push_ms %.2f' % d['breakdown_ms_per_step']['push_ms'])"
done
