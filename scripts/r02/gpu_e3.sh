#!/bin/bash
# Round 2, call E3 (2 GPUs), final build: the NCCL all-reduce test and bench.py at N = 2
# (weak line, strong-scaling pass, in-bench all-reduce value check), both arms under torchrun.
mkdir -p gpurun_out
T=r02e3
nvidia-smi -L
python -m pytest tests -m gpu -q -k "two_gpus or nccl" > gpurun_out/${T}_pytest_two_gpus.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_two_gpus.log
tail -5 gpurun_out/${T}_pytest_two_gpus.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/${T}_bench_c1_n2.json 2> gpurun_out/${T}_bench_c1_n2.err
python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_c1_n2.json').read().strip().splitlines()[-1]);print('N=2 value %.4g e2e %.4g' % (d['value'], d['e2e']['value']), d.get('allreduce_check'), d.get('strong_scaling'))" || tail -5 gpurun_out/${T}_bench_c1_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/${T}_ref_c1_n2.json 2> gpurun_out/${T}_ref_c1_n2.err
python -c "
import json;d=json.loads(open('gpurun_out/${T}_ref_c1_n2.json').read().strip().splitlines()[-1]);print('reference arm under torchrun: value %.4g cores %d' % (d['value'], d['cpu_baseline']['cores']))" || tail -5 gpurun_out/${T}_ref_c1_n2.err
