mkdir -p gpurun_out
for pad in 0 40000 60000 100000; do
  GPAT_PUSH_SMEM_PAD=$pad timeout 300 python bench.py --workload c5 --grid 256 --nptl 2000000 --no-cpu-baseline --steps 4 --warmup 2 > gpurun_out/occ_c5_$pad.json 2>> gpurun_out/occ_c5.err
  python -c "
import json;d=json.load(open('gpurun_out/occ_c5_$pad.json'));print('c5 pad $pad value %.4g push_ms %.2f' % (d['value'], d['breakdown_ms_per_step']['push_ms']))"
done
