#!/bin/bash
# usage (GPU box): [PARITY=1] tools/variant_bench.sh <kernel-regex> v1.so v2.so ...   -- per-variant ncu launch times of one
# kernel (variants are complete libpslot_orb.so builds under tmp_variants/; the in-tree library is restored afterwards);
# PARITY=1 also runs the extractor / stereo parity tests against each variant
K=$1; shift
cp pointslot_b200/lib/libpslot_orb.so /tmp/orig.so
for v in "$@"; do
  cp tmp_variants/$v pointslot_b200/lib/libpslot_orb.so
  if [ -n "$PARITY" ]; then echo "$v parity: $(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_property.py -q -x 2>&1 | tail -1)"; fi
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:$K -c 40 --csv --log-file gpurun_out/var_$v.csv \
      python tools/profile_driver.py --batch 125 --submits 2 > /dev/null 2>&1
  echo "$v: $(python tools/launch_summary.py gpurun_out/var_$v.csv | tr '\n' ';')"
done
cp /tmp/orig.so pointslot_b200/lib/libpslot_orb.so
