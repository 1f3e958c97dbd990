#!/bin/bash
# usage (GPU box): tools/variant_bench.sh <kernel-regex> v1.so v2.so ...   -- per-variant ncu launch times of one kernel
# (variants are complete libpslot_orb.so builds under tmp_variants/; the in-tree library is restored afterwards)
K=$1; shift
cp pointslot_b200/lib/libpslot_orb.so /tmp/orig.so
for v in "$@"; do
  cp tmp_variants/$v pointslot_b200/lib/libpslot_orb.so
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:$K -c 40 --csv --log-file gpurun_out/var_$v.csv \
      python tools/profile_driver.py --batch 125 --submits 2 > /dev/null 2>&1
  echo "$v: $(python tools/launch_summary.py gpurun_out/var_$v.csv | tr '\n' ';')"
done
cp /tmp/orig.so pointslot_b200/lib/libpslot_orb.so
