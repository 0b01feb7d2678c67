python -c "
import torch; p=torch.cuda.get_device_properties(0); print('L2', p.L2_cache_size, p.L2_cache_size/2**20, 'SMs', p.multi_processor_count)"
timeout 300 python bench.py --workload c5 --grid 256 --nptl 2000000 --no-cpu-baseline --steps 4 --warmup 2 > gpurun_out/l2_c5.json 2> gpurun_out/l2_c5.err
python -c "
import json;d=json.load(open('gpurun_out/l2_c5.json'));print('c5 value %.4g push_ms %.2f' % (d['value'], d['breakdown_ms_per_step']['push_ms']))"
timeout 300 python bench.py --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/l2_c1.json 2> gpurun_out/l2_c1.err
python -c "
import json;d=json.load(open('gpurun_out/l2_c1.json'));print('c1 value %.4g push_ms %.2f' % (d['value'], d['breakdown_ms_per_step']['push_ms']))"
timeout 300 python bench.py --workload c4 --nptl 300000 --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/l2_c4.json 2> gpurun_out/l2_c4.err
python -c "
import json;d=json.load(open('gpurun_out/l2_c4.json'));print('c4 value %.4g push_ms %.2f' % (d['value'], d['breakdown_ms_per_step']['push_ms']))"
python -m pytest tests -m gpu -x -q -k "c5 or 3d" 2>&1 | tail -2
