python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/q2_pytest.log
for lib in bonnie-32_b200/libb32raster.so build/v_ordgroup1.so build/v_ordgroup4.so; do
  echo "== $lib" >> gpurun_out/q2_perf.txt
  B32_LIB=$PWD/$lib python tools/perf_scenes.py xray transparent big_tri float mixed >> gpurun_out/q2_perf.txt 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:k_fill_ordered -c 1 -f -o gpurun_out/q2_ordered python tools/ncu_probe.py ordered > gpurun_out/q2_ncu.log 2>&1
cat gpurun_out/q2_pytest.log; cat gpurun_out/q2_perf.txt
