mkdir -p gpurun_out
for v in 0 1; do
  echo "== variant $v" >> gpurun_out/job1.log
  GPAT_PUSH_VARIANT=$v python -m pytest tests -m gpu -x -q 2>&1 | tail -8 >> gpurun_out/job1.log
  GPAT_PUSH_VARIANT=$v python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v$v.json 2>> gpurun_out/job1.log
done
GPAT_PUSH_VARIANT=1 ncu --set full --clock-control none --import-source on -k regex:push_kernel -s 1 -c 1 -o gpurun_out/prof_coop python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_coop.log 2>&1
cat gpurun_out/job1.log; for v in 0 1; do python -c "
import json;d=json.load(open('gpurun_out/bench_v$v.json'));print($v, d['value'], d['roofline']['frac'], d['breakdown_ms_per_step'])"; done
