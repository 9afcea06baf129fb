#!/bin/bash
# Round-2 session E (1 GPU): top-K cached arg-max + self-stamped entries + direct encode output on hardware; ncu evidence.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "### probe_train 100 MB zipf"; timeout 300 python tools/probe_train.py zipf 32000 100e6 2>&1 | tail -1
echo "### probe_train threads 512"; YTTM_LOOP_THREADS=512 timeout 300 python tools/probe_train.py zipf 32000 100e6 2>&1 | tail -1
echo "### probe_train readme"; timeout 300 python tools/probe_train.py readme 2>&1 | tail -1
echo "### pytest -m gpu"; timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "### ab_encode"; timeout 300 python tools/ab_encode.py 1000000 5 gpurun_out/r02e_ab_encode.json 2>&1 | tail -3
echo "### bench"; timeout 900 python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02e_bench.err
echo "### ncu launch list (encode only)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/r02e_launches.csv python bench.py --steps 2 --warmup 3 --no-train-legs --no-cpu-baseline > /dev/null 2>&1; echo "rc=$?"
echo "### ncu --set full: encode kernels"; timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'find_words_vec|dedup_words|encode_rep_words|sentence_ids|emit_ids' -s 10 -c 5 -f -o gpurun_out/r02e_prof_encode \
    python bench.py --steps 2 --warmup 3 --no-train-legs --no-cpu-baseline > /dev/null 2>&1; echo "rc=$?"
echo "### ncu --set full: byte passes (100 MB zipf)"; timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'char_hist_kernel|word_insert_kernel' -c 2 -f -o gpurun_out/r02e_prof_front \
    python tools/probe_train.py zipf 32000 100e6 > /dev/null 2>&1; echo "rc=$?"
echo "### ncu --set full: merge loop RESIDENT (100 MB zipf, the long launch)"; timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'^merge_loop_kernel' -s 2 -c 1 -f -o gpurun_out/r02e_prof_merge_loop_resident \
    python tools/probe_train.py zipf 32000 100e6 > /dev/null 2>&1; echo "rc=$?"
} > gpurun_out/r02_sessionE.log 2>&1
tail -70 gpurun_out/r02_sessionE.log
