# usage: bash scripts/gpu_profile.sh <out tag> [bench args...]   -> gpurun_out/prof_<tag>.ncu-rep (+ launches_<tag>.csv)
tag=$1; shift
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:push_kernel -s 1 -c 1 -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/ncu_$tag.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/ncu_l_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log | cut -c1-300
