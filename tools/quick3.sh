python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/q3_pytest.log
rm -f gpurun_out/q3_perf.txt
for lib in bonnie-32_b200/libb32raster.so build/v_ordsub1.so build/v_ordsub4.so; do
  echo "== $lib" >> gpurun_out/q3_perf.txt
  B32_LIB=$PWD/$lib python tools/perf_scenes.py xray transparent mixed >> gpurun_out/q3_perf.txt 2>&1
done
python tools/perf_scenes.py c4_1M c4_1920 c4_640 c4_100000 sky >> gpurun_out/q3_perf.txt 2>&1
python tools/fill_stats_1m.py > gpurun_out/q3_stats1m.txt 2>&1
cat gpurun_out/q3_pytest.log; cat gpurun_out/q3_perf.txt; cat gpurun_out/q3_stats1m.txt
