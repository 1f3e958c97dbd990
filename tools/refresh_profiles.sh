#!/bin/bash
# usage (GPU box, via gpurun): tools/refresh_profiles.sh   -- everything profiles/ is built from, into gpurun_out/
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --no-determinism > gpurun_out/bench_under_ncu.log 2>&1
for k in k_fast_cells k_describe k_resize_level k_quadtree k_stereo_match k_gray_repitch; do
  tools/prof.sh $k prof_$k
done
# the cv::ORB row: launch list of two 64-image batched calls (pageable host images: four chunks of 16) and one full capture
# per kernel
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cvorb.csv \
    python tools/cvorb_profile.py 64 > gpurun_out/cvorb_under_ncu.log 2>&1
for k in k_cv_fast_score k_cv_nms_collect k_cv_harris k_cv_describe k_cv_resize_exact k_cv_select; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/prof_$k \
      python tools/cvorb_profile.py 64 > gpurun_out/prof_$k.log 2>&1
done
python tools/cvorb_bench.py 128 > gpurun_out/cvorb_bench.json 2> gpurun_out/cvorb_bench.err
python bench.py > gpurun_out/bench_full.log 2>&1; tail -1 gpurun_out/bench_full.log > gpurun_out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log > gpurun_out/bench_reference.json
