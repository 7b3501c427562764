ncu --set full --clock-control none --import-source on -k regex:"k_setup|k_fill_opaque" -s 4 -c 8 -f -o gpurun_out/q10_flushed python tools/ncu_c4.py > gpurun_out/q10_ncu.log 2>&1
tail -2 gpurun_out/q10_ncu.log
