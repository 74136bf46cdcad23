#!/bin/bash
# same-box A/B: conv_tc2 of commit 2c5f38d (4 epilogue warps, per-thread scattered stores) vs the current kernel
cp dynavsr_b200/libdvsr_b200.so /tmp/lib_default.so
for v in old new old new; do
  if [ $v = old ]; then cp tools/variants/lib_oldtc2.so dynavsr_b200/libdvsr_b200.so; else cp /tmp/lib_default.so dynavsr_b200/libdvsr_b200.so; fi
  for prec in bf16x3 bf16; do for shape in "5 176 320" "1 44 80"; do
    echo -n "$v conv $shape: "; timeout 120 python tools/one_conv.py $shape 64 64 3 --precision $prec 2>&1 | tail -1
  done; done
done
for v in old new; do
  if [ $v = old ]; then cp tools/variants/lib_oldtc2.so dynavsr_b200/libdvsr_b200.so; else cp /tmp/lib_default.so dynavsr_b200/libdvsr_b200.so; fi
  timeout 300 python bench.py --steps 36 --warmup 6 --no-reference-cuda --no-cpu-baseline --no-parity --no-roofline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$v: value %.2f e2e %.2f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
  timeout 300 python bench.py --steps 36 --warmup 6 --workload infer --no-reference-cuda --no-cpu-baseline --no-parity --no-roofline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$v infer: value %.2f e2e %.2f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
cp /tmp/lib_default.so dynavsr_b200/libdvsr_b200.so
