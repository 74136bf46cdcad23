#!/bin/bash
# staged DCN window margin x operand-ring depth variants (tools/variants/lib_m<margin>_s<stages>.so swapped in for the library)
cp dynavsr_b200/libdvsr_b200.so /tmp/lib_default.so
for v in m4_s2 m5_s2 m5_s3 m3_s2 m6_s2; do
  cp tools/variants/lib_$v.so dynavsr_b200/libdvsr_b200.so
  for s in 1.0 1.5 3.0; do echo -n "$v offset std $s: "; timeout 60 python tools/one_dcn.py 5 176 320 --offset-std $s 2>&1 | tail -1; done
done
cp /tmp/lib_default.so dynavsr_b200/libdvsr_b200.so
