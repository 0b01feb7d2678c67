mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:push_kernel -s 1 -c 1 -o gpurun_out/prof_c1e python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c1e.log 2>&1
tail -1 gpurun_out/ncu_c1e.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c1e.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l_c1e.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:push_kernel -s 1 -c 1 -o gpurun_out/prof_c4 python bench.py --workload c4 --grid 1024 --nptl 150000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c4.log 2>&1
tail -1 gpurun_out/ncu_c4.log | cut -c1-200
