set -x
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/probe_launches.csv python tools/ncu_probe.py > gpurun_out/probe.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fill_ordered -c 2 -f -o gpurun_out/probe_ordered python tools/ncu_probe.py ordered >> gpurun_out/probe.log 2>&1
tail -3 gpurun_out/probe.log
