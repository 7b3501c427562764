set -x
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r02_bench_reference.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 60 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu --inflight 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 120 --csv --log-file gpurun_out/r02_launches_inflight4.csv python bench.py --steps 60 --warmup 3 --no-cpu --inflight 4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_setup|k_fill_opaque" -s 4 -c 8 -f -o gpurun_out/r02_flushed python tools/ncu_c4.py > gpurun_out/r02_ncu.log 2>&1
ncu --set full --clock-control none --cache-control none -k regex:"k_setup|k_fill_opaque" -s 4 -c 8 -f -o gpurun_out/r02_steady python tools/ncu_c4.py >> gpurun_out/r02_ncu.log 2>&1
python tools/perf_scenes.py > gpurun_out/r02_perf_scenes.txt 2>&1
build/call_overhead >> gpurun_out/r02_perf_scenes.txt 2>&1
python tools/fill_stats.py > gpurun_out/r02_fill_stats.txt 2>&1
python tools/fill_stats_1m.py > gpurun_out/r02_fill_stats_1m.txt 2>&1
python tests/checks/perf_game_frame.py > gpurun_out/r02_perf_game.txt 2>&1
bash tools/sanitize.sh
cat gpurun_out/r02_pytest.log; tail -3 gpurun_out/r02_bench.err
