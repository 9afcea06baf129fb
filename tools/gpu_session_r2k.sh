#!/bin/bash
# Round-2 session K (1 GPU, short): front-end of the config-5 (multilingual) corpus — why is its word table 5x slower per byte?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "### probe_train multilingual 125 MB"; timeout 600 python tools/probe_train.py multilingual 64000 125e6 2>&1 | tail -2 | cut -c1-900
echo "### ncu byte passes (multilingual)"; timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'char_hist_kernel|word_insert_kernel|pair_hist_kernel|word_compact|word_tokens' -c 6 -f -o gpurun_out/r02k_prof_front_multi \
    python tools/probe_train.py multilingual 64000 125e6 > /dev/null 2>&1; echo "rc=$?"
ncu -i gpurun_out/r02k_prof_front_multi.ncu-rep --page raw --csv > gpurun_out/r02k_prof_front_multi.raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02k_prof_front_multi.raw.csv | cut -c1-200
} > gpurun_out/r02_sessionK.log 2>&1
cat gpurun_out/r02_sessionK.log
