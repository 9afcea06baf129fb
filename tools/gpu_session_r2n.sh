#!/bin/bash
# Round-2 final 1-GPU session: GPU test suite, bench line, reference arm, loop probes (no ncu: see session J).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "### pytest -m gpu"; timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "### bench"; timeout 1200 python bench.py > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; echo "bench rc=$?"; tail -c 200 gpurun_out/r02n_bench.err
echo "### bench --impl reference"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02n_bench_reference.json 2> gpurun_out/r02n_bench_reference.err; echo "rc=$?"
echo "### probe_train 100 MB zipf"; timeout 300 python tools/probe_train.py zipf 32000 100e6 2>&1 | tail -1 | cut -c1-700
echo "### probe_train readme"; timeout 300 python tools/probe_train.py readme 2>&1 | tail -1 | cut -c1-700
echo "### YTTM_DBG=16"; YTTM_DBG=16 timeout 300 python tools/probe_train.py zipf 32000 100e6 2>&1 | grep DBG16 | head -11
} > gpurun_out/r02_sessionN.log 2>&1
cut -c1-700 gpurun_out/r02_sessionN.log | tail -40
