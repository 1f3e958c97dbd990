#!/bin/bash
# usage (GPU box, via gpurun): tools/refresh_profiles.sh   -- everything profiles/ is built from, into gpurun_out/
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --no-determinism > gpurun_out/bench_under_ncu.log 2>&1
for k in k_fast_cells k_describe k_resize_level k_quadtree k_stereo_match k_gray_repitch; do
  tools/prof.sh $k prof_$k
done
python bench.py > gpurun_out/bench_full.log 2>&1; tail -1 gpurun_out/bench_full.log > gpurun_out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log > gpurun_out/bench_reference.json
