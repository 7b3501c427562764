#!/bin/bash
# usage (on the GPU box): tools/bench_variants.sh build/vA.so build/vB.so ...   -> value / e2e / kernel times per build
for lib in "$@"; do
  for inf in 1 4; do
    B32_LIB=$PWD/$lib python bench.py --no-cpu --inflight $inf 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib inflight=$inf value %.0f e2e %.0f kernels %s exact %s' % (d['value'], d['e2e']['value'], {k: round(v*1000,1) for k,v in d['roofline']['kernel_ms'].items() if v}, d['bit_exact_vs_golden']))"
  done
done
