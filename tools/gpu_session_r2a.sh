#!/bin/bash
# Round-2 session A (one gpurun call): A/B of the env-gated encode kernels and merge-loop knobs written at the end of
# round 1, plus racecheck / synccheck of the tiny sanitizer workload.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "### ab_encode"; timeout 300 python tools/ab_encode.py 1000000 7 gpurun_out/r02_ab_encode.json 2>&1 | tail -3
echo "### ab_long_words"; timeout 180 python tools/ab_long_words.py 16384 8 2>/dev/null | tail -20
echo "### ab_train"; timeout 500 bash tools/ab_train.sh zipf 32000 100e6
echo "### sanitizers"
for t in racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $t --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/r02_sanitizer_$t.log 2>&1; echo "$t rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small:" gpurun_out/r02_sanitizer_$t.log | tail -3
done
} > gpurun_out/r02_sessionA.log 2>&1
tail -30 gpurun_out/r02_sessionA.log
