#!/bin/bash
# Round 2, call Q (8 GPUs): BASELINE config[4] as stated (512^3 field per GPU, 1e9 particles over 8 GPUs) with the final 3-D kernel.
mkdir -p gpurun_out
T=r02q
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 1200 $TR --master-port 29533 bench.py --workload c5 --nptl 125000000 --gpus 8 --steps 2 --warmup 1 --no-membw --no-cpu-baseline > gpurun_out/${T}_full_c5_n8.json 2> gpurun_out/${T}_full_c5_n8.err
python -c "
import json;d=json.loads(open('gpurun_out/${T}_full_c5_n8.json').read().strip().splitlines()[-1]);print('C5 x8 value %.4g e2e %.4g ms/step %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d.get('allreduce_check'), d['clocks'])" || tail -5 gpurun_out/${T}_full_c5_n8.err
