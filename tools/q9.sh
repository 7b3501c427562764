python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/q9_pytest.log
for lib in bonnie-32_b200/libb32raster.so; do
  echo "== $lib" >> gpurun_out/q9.txt
  B32_LIB=$PWD/$lib ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:"k_setup|k_fill_opaque" -s 4 -c 8 --csv python tools/ncu_c4.py 2>/dev/null | grep -v "^==" | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
for r in rows[1:]:
    if len(r)>5: print(r[4][:40].replace(chr(10),' '), r[-3], r[-1])" >> gpurun_out/q9.txt
  for rep in 1 2; do B32_LIB=$PWD/$lib python bench.py --no-cpu --steps 400 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.0f e2e %.0f h2d %d' % (d['value'], d['e2e']['value'], d['e2e']['h2d_bytes_per_step']))" >> gpurun_out/q9.txt; done
done
cat gpurun_out/q9_pytest.log gpurun_out/q9.txt
