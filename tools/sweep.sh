#!/bin/bash
# usage (GPU box): tools/sweep.sh  -> value per build variant / frames in flight
one() { B32_LIB=$1 python bench.py --no-cpu --steps 300 --inflight $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-28s inflight=%-2s value %6.0f  e2e %5.0f  game %5.0f  kernels %s exact %s' % ('$1'.split('/')[-1], '$2', d['value'], d['e2e']['value'], d['e2e']['resident_geometry']['value'], {k: round(v*1000,1) for k,v in d['roofline']['kernel_ms'].items()}, d['bit_exact_vs_golden']))"; }
main=$PWD/bonnie-32_b200/libb32raster.so
for inf in 1 2 4 6 8 12; do one $main $inf; done
for v in build/v_*.so; do one $PWD/$v 4; done
