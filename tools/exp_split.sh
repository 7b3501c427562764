#!/bin/bash
# usage (GPU box): tools/exp_split.sh -> fused front end vs B32_SPLIT_TRANSFORM=1 (k_transform + k_setup): parity, value, blocking-call cost
one() { env $1 python bench.py --no-cpu --steps $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-24s steps %-4s value %6.0f  e2e %5.0f  game %5.0f  sync_call_ms %.4f  launches %s kernels %s exact %s' % ('$1', '$2', d['value'], d['e2e']['value'], d['e2e']['resident_geometry']['value'], d['sync_call_ms'], d['gpu_launches'], {k: round(v*1000,1) for k,v in d['roofline']['kernel_ms'].items()}, d['bit_exact_vs_golden']))"; }
for rep in 1 2; do
  for steps in 20 400; do
    one "B32_X=0" $steps
    one "B32_SPLIT_TRANSFORM=1" $steps
  done
done
echo "== native blocking call"; build/call_overhead 2>&1 | tail -6
echo "== native blocking call, split"; B32_SPLIT_TRANSFORM=1 build/call_overhead 2>&1 | tail -6
echo "== parity with the split front end"
B32_SPLIT_TRANSFORM=1 python -m pytest tests/test_gpu_parity.py tests/test_c3_scenes.py -m gpu -x -q 2>&1 | tail -4
