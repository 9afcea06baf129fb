#!/bin/bash
# Round-2 session J (1 GPU, the full one): GPU test suite, bench line, ncu launch list + full captures of the encode
# kernels, the byte passes and the RESIDENT merge loop.  Everything under gpurun_out/r02j_*.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "### probe_train 100 MB zipf"; timeout 300 python tools/probe_train.py zipf 32000 100e6 2>&1 | tail -1
echo "### probe_train readme"; timeout 300 python tools/probe_train.py readme 2>&1 | tail -1
echo "### YTTM_DBG=16"; YTTM_DBG=16 timeout 300 python tools/probe_train.py zipf 32000 100e6 2>&1 | grep DBG16 | head -12
echo "### pytest -m gpu"; timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "### bench"; timeout 1200 python bench.py > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02j_bench.err
echo "### bench --impl reference"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02j_bench_reference.json 2> gpurun_out/r02j_bench_reference.err; echo "rc=$?"
echo "### ncu launch list (encode only)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/r02j_launches.csv python bench.py --steps 2 --warmup 3 --no-train-legs --no-cpu-baseline > /dev/null 2>&1; echo "rc=$?"
echo "### ncu --set full: encode kernels"; timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'find_words_vec|dedup_words|encode_rep_words|sentence_ids|emit_ids' -s 10 -c 5 -f -o gpurun_out/r02j_prof_encode \
    python bench.py --steps 2 --warmup 3 --no-train-legs --no-cpu-baseline > /dev/null 2>&1; echo "rc=$?"
echo "### ncu --set full: byte passes (100 MB zipf)"; timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'char_hist_kernel|word_insert_kernel|pair_hist_kernel' -c 3 -f -o gpurun_out/r02j_prof_front \
    python tools/probe_train.py zipf 32000 100e6 > /dev/null 2>&1; echo "rc=$?"
echo "### ncu --set full: merge loop RESIDENT (100 MB zipf, the last launch)"; timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'^merge_loop_kernel' -s 4 -c 1 -f -o gpurun_out/r02j_prof_merge_loop \
    python tools/probe_train.py zipf 32000 100e6 > /dev/null 2>&1; echo "rc=$?"
for f in prof_encode prof_front prof_merge_loop; do
  ncu -i gpurun_out/r02j_$f.ncu-rep --page raw --csv > gpurun_out/r02j_$f.raw.csv 2>/dev/null
done
ls -la gpurun_out/ | grep r02j
} > gpurun_out/r02_sessionJ.log 2>&1
tail -60 gpurun_out/r02_sessionJ.log | cut -c1-600
