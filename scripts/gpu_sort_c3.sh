for s in 0 1 0 1; do
GPAT_PUSH_SORT=$s timeout 300 python bench.py --workload c3 --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/c3_sort$s.json 2>> gpurun_out/c3_sort.err
python -c "
import json;d=json.load(open('gpurun_out/c3_sort$s.json'));print('c3 sort=$s value %.4g push_ms %.2f' % (d['value'], d['breakdown_ms_per_step']['push_ms']))"
done
