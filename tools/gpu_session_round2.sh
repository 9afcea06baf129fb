#!/bin/bash
# First GPU session of the next round, one gpurun call (≈ 20–30 GPU-minutes with the sanitizer passes and the six ncu captures;
# drop sections to fit the budget - every section is independent):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_session_round2.sh'
# 1. the GPU tests (incl. the reference's own stress test binary and the byte-identical model file test),
# 2. A/B of the experimental encode kernels and of the merge-loop geometry,
# 3. compute-sanitizer memcheck / racecheck / synccheck on a tiny workload,
# 4. ncu: launch list of bench.py and --set full captures of encode_words (default, bucketed) and of the dedup / vector-find kernels.
# Everything lands in gpurun_out/ ; copy what is to be judged into profiles/ (r02_*).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "### pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "### ab_encode"; timeout 300 python tools/ab_encode.py 1000000 7 2>/dev/null | tail -110
echo "### ab_long_words"; timeout 180 python tools/ab_long_words.py 16384 8 2>/dev/null | tail -20
echo "### ab_train"; timeout 400 bash tools/ab_train.sh zipf 32000 100e6
echo "### compute-sanitizer (tools/sanitize_small.py: tiny train + encode incl. the experimental variants, checked against the oracle)"
for t in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $t --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/r02_sanitizer_$t.log 2>&1; echo "$t rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small:" gpurun_out/r02_sanitizer_$t.log | tail -3
done
echo "### bench"; timeout 600 python bench.py > gpurun_out/r02_bench_first.json 2> gpurun_out/r02_bench_first.err; echo "bench rc=$?"
echo "### ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --scan-tokens 67108864 > /dev/null 2>&1; echo "rc=$?"
echo "### ncu encode_words (default)"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:encode_words -s 2 -c 1 -f \
    -o gpurun_out/r02_prof_encode_words python tools/ab_encode.py 1000000 1 > /dev/null 2>&1; echo "rc=$?"
echo "### ncu encode_words (bucketed)"; YTTM_ENC_BUCKETED=1 timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:bucketed -s 2 -c 1 -f -o gpurun_out/r02_prof_encode_words_bucketed python tools/ab_encode.py 1000000 1 > /dev/null 2>&1; echo "rc=$?"
echo "### ncu find_vec + dedup kernels (ab_encode sets the variants itself: 21 matching launches precede the dedup+find_vec variant)"; timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:'find_words_vec|dedup_words|encode_rep_words|copy_word_ids' -s 21 -c 4 -f -o gpurun_out/r02_prof_encode_dedup python tools/ab_encode.py 1000000 1 > /dev/null 2>&1; echo "rc=$?"
echo "### ncu merge loop on the 100 MB training (RESIDENT): default and wide-probe"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'^merge_loop_kernel' -s 1 -c 1 -f \
    -o gpurun_out/r02_prof_merge_loop_resident python tools/probe_train.py zipf 32000 100e6 > /dev/null 2>&1; echo "rc=$?"
YTTM_LOOP_WIDEPROBE=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:merge_loop_wide -s 1 -c 1 -f \
    -o gpurun_out/r02_prof_merge_loop_resident_wide python tools/probe_train.py zipf 32000 100e6 > /dev/null 2>&1; echo "rc=$?"
} > gpurun_out/r02_session1.log 2>&1
tail -5 gpurun_out/r02_session1.log
