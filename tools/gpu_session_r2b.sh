#!/bin/bash
# Round-2 session B (one gpurun call, 1 GPU): the owner-computes merge loop and the new encode defaults on hardware.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "### probe_train 100 MB zipf (new merge loop)"; timeout 300 python tools/probe_train.py zipf 32000 100e6 2>&1 | tail -2
echo "### probe_train YTTM_DBG=8"; YTTM_DBG=8 timeout 300 python tools/probe_train.py zipf 32000 100e6 2>&1 | tail -1
echo "### probe_train threads 512"; YTTM_LOOP_THREADS=512 timeout 300 python tools/probe_train.py zipf 32000 100e6 2>&1 | tail -1
echo "### probe_train readme"; timeout 300 python tools/probe_train.py readme 2>&1 | tail -1
echo "### pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "### ab_encode"; timeout 300 python tools/ab_encode.py 1000000 5 gpurun_out/r02_ab_encode_b.json 2>&1 | tail -3
echo "### sanitizers"
for t in racecheck synccheck memcheck; do
  timeout 500 compute-sanitizer --tool $t --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/r02b_sanitizer_$t.log 2>&1; echo "$t rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small:" gpurun_out/r02b_sanitizer_$t.log | tail -3
done
echo "### bench"; timeout 900 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02b_bench.err
} > gpurun_out/r02_sessionB.log 2>&1
tail -60 gpurun_out/r02_sessionB.log
