mkdir -p gpurun_out
for g in 1 0 1 0; do
GPAT_PUSH_GENERIC=$g timeout 300 python bench.py --workload c5 --grid 256 --nptl 2000000 --no-cpu-baseline --steps 6 --warmup 2 > gpurun_out/bench8_c5_g$g.json 2> gpurun_out/bench8_c5.err
python -c "
import json;d=json.load(open('gpurun_out/bench8_c5_g$g.json'));print('c5 generic=$g value %.4g push_ms %.2f' % (d['value'], d['breakdown_ms_per_step']['push_ms']))"
done
for g in 1 0; do
GPAT_PUSH_GENERIC=$g timeout 400 python bench.py --workload c4 --nptl 300000 --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/bench8_c4_g$g.json 2> gpurun_out/bench8_c4.err
python -c "
import json;d=json.load(open('gpurun_out/bench8_c4_g$g.json'));print('c4 generic=$g value %.4g push_ms %.2f' % (d['value'], d['breakdown_ms_per_step']['push_ms']))"
done
