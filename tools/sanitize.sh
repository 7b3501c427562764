# usage (GPU box): bash tools/sanitize.sh   -> gpurun_out/r02_sanitizer.txt (memcheck, racecheck, synccheck over tests/checks/sanitize_run.py)
rm -f gpurun_out/r02_sanitizer.txt
for t in memcheck racecheck synccheck; do echo "== $t" >> gpurun_out/r02_sanitizer.txt; timeout 900 compute-sanitizer --tool $t python tests/checks/sanitize_run.py 2>&1 | tail -70 >> gpurun_out/r02_sanitizer.txt; done
grep -n "MISMATCH\|mismatches\|ERROR SUMMARY" gpurun_out/r02_sanitizer.txt
