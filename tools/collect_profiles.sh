#!/bin/bash
# usage: tools/collect_profiles.sh   (after tools/r02_measure.sh ran under gpurun: copies its outputs from gpurun_out/ into profiles/)
set -e
cd "$(dirname "$0")/.."
o=gpurun_out
cp $o/r02_bench.json $o/r02_bench_reference.json $o/r02_launches.csv $o/r02_launches_inflight4.csv $o/r02_perf_scenes.txt \
   $o/r02_fill_stats.txt $o/r02_fill_stats_1m.txt $o/r02_perf_game.txt profiles/
cp $o/r02_sanitizer.txt profiles/r02_compute_sanitizer.txt
cp $o/r02_pytest.log profiles/r02_pytest_gpu.txt
python tools/ncu_kernels_json.py $o/r02_flushed.ncu-rep $o/r02_steady.ncu-rep > profiles/r02_ncu_kernels.json
python tools/ncu_summary.py $o/r02_flushed.ncu-rep > profiles/r02_ncu_summary.txt
{ echo "## k_setup (C4 frame, cache-flushed capture)"; python tools/ncu_lines.py $o/r02_flushed.ncu-rep k_setup 0 30; echo
  echo "## k_fill_opaque, OpSparse shape (256 threads) - the shape the enqueued frames of bench.py run"; python tools/ncu_lines.py $o/r02_flushed.ncu-rep k_fill_opaque 1 40; echo
  echo "## k_fill_opaque, OpDense shape (512 threads) - the blocking-call shape"; python tools/ncu_lines.py $o/r02_flushed.ncu-rep k_fill_opaque 0 30; } > profiles/r02_ncu_hot_lines.txt
bash tools/sass_opcodes.sh > profiles/r02_sass_opcodes.txt 2>&1 || true
