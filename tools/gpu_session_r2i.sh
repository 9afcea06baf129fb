#!/bin/bash
# Round-2 session I (1 GPU): ncu source-level capture of the RESIDENT merge loop (where do the warps stall?).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "### probe"; timeout 300 python tools/probe_train.py zipf 32000 100e6 2>&1 | tail -1 | cut -c1-400
echo "### ncu --set full: merge loop RESIDENT (100 MB zipf, last launch)"; timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'^merge_loop_kernel' -s 4 -c 1 -f -o gpurun_out/r02i_prof_merge_loop \
    python tools/probe_train.py zipf 32000 100e6 > /dev/null 2>&1; echo "rc=$?"
ls -la gpurun_out/r02i_prof_merge_loop.ncu-rep
} > gpurun_out/r02_sessionI.log 2>&1
cat gpurun_out/r02_sessionI.log
