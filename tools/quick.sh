set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/q_pytest.log
python tools/perf_scenes.py > gpurun_out/q_perf_scenes.txt 2>&1
build/call_overhead >> gpurun_out/q_perf_scenes.txt 2>&1
python bench.py --no-cpu > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
cat gpurun_out/q_pytest.log; head -20 gpurun_out/q_perf_scenes.txt
