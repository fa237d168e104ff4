#!/bin/bash
# bench several kernel variants back to back: tools/bench_variants.sh <workload> v1 v2 ...
W=$1; shift
for v in "$@"; do
  python bench.py --steps 10 --no-cpu --workload $W --variant $v 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-14s %8.1f Mpix/s  %7.3f ms  e2e %8.1f  regs %3d  ctas/sm %d  %s' % ('$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['regs'], d['config']['ctas_per_sm'], d['clocks']['reasons']))"
done
