# usage: bash scripts/gpu_variants.sh [tag ...]   (tags of csrc/libgpat_cuda.<tag>.so; "" = default)
mkdir -p gpurun_out
LOG=gpurun_out/variants.log
: > $LOG
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $LOG
run() {  # tag, extra env
  tag=$1; lib=stochastic_parker_b200/csrc/libgpat_cuda${tag:+.$tag}.so
  GPAT_LIB=$PWD/$lib python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag:-default}.json 2>> $LOG
  python - <<PY >> $LOG
import json
d=json.load(open('gpurun_out/bench_${tag:-default}.json'))
print('${tag:-default}', '%.4g steps/s' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'push_ms %.2f' % d['breakdown_ms_per_step']['push_ms'])
PY
}
run ""
for t in "$@"; do
  GPAT_LIB=$PWD/stochastic_parker_b200/csrc/libgpat_cuda.$t.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "step_parity or interval_parity_fast" 2>&1 | tail -2 >> $LOG
  run $t
done
#ncu --set full --clock-control none --import-source on -k regex:push_kernel -s 1 -c 1 -o gpurun_out/prof_default python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_default.log 2>&1
cat $LOG
