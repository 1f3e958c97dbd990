#!/bin/bash
# usage (on the GPU box, via gpurun): tools/prof.sh <kernel-regex> <out-name> [driver args]
# one ncu --set full capture of the first matching launch of the bench sub-batch workload
set -e
K=$1; N=$2; shift 2
ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o gpurun_out/$N \
    timeout 600 python tools/profile_driver.py --batch 125 --submits 1 "$@" > gpurun_out/$N.log 2>&1
