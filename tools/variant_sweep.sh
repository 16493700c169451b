#!/bin/bash
# usage: tools/variant_sweep.sh <workload> <variant> [<variant> ...]   ("base" = the shipped library)
# prints kernel ms / value / roofline fraction of bench.py for every variant library built by tools/build_variant.py
w=$1; shift
for v in "$@"; do
  if [ "$v" = base ]; then unset DFX_LIB; else export DFX_LIB=diffrax_b200/lib/variants/libdiffrax_b200_$v.so; fi
  timeout 200 python bench.py --workload $w --steps 20 --warmup 3 --no-extras --cpu-sample 4096 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w $v kernel_ms %.4f ms %.4f value %.4e frac %.4f failed %d' % (d['kernel_ms_per_step'], d['ms_per_step'], d['value'], d['roofline']['frac'], d['config']['failed_trajectories']))"
done
