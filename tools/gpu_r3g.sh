#!/bin/bash
# call 3G: full racecheck report of the tile kernel (the 8 warnings of call 3Z)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 40 python tools/sanitize_target.py > gpurun_out/racecheck_3g_full.log 2>&1
grep -c "Race reported\|hazard" gpurun_out/racecheck_3g_full.log; grep "RACECHECK SUMMARY" gpurun_out/racecheck_3g_full.log
head -c 6000 gpurun_out/racecheck_3g_full.log
