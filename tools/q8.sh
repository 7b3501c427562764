python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/q8_pytest.log
python tools/perf_scenes.py c4_100000 c1_ c2_1000_tris_64x64_idx8 > gpurun_out/q8_perf.txt 2>&1
build/call_overhead >> gpurun_out/q8_perf.txt 2>&1
for inf in 2 3 4 6 8; do for rep in 1 2 3; do python bench.py --steps 20 --warmup 5 --no-cpu --inflight $inf 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('inflight $inf steps 20 value %.0f e2e %.0f' % (d['value'], d['e2e']['value']))" >> gpurun_out/q8_perf.txt; done; done
cat gpurun_out/q8_pytest.log gpurun_out/q8_perf.txt
