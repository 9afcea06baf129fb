#!/bin/bash
# Round-2 session L (1 GPU): bench line after the sketch / table-size changes, config-5 front end at 1.25 GB, loop probe.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "### probe_train 100 MB zipf"; timeout 300 python tools/probe_train.py zipf 32000 100e6 2>&1 | tail -1 | cut -c1-700
echo "### probe_train readme"; timeout 300 python tools/probe_train.py readme 2>&1 | tail -1 | cut -c1-700
echo "### bench"; timeout 1200 python bench.py > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err; echo "bench rc=$?"; tail -c 200 gpurun_out/r02p_bench.err
echo "### train parity tests"; timeout 900 python -m pytest tests/test_train_gpu.py tests/test_train_scale_gpu.py -x -q -m gpu 2>&1 | tail -3
} > gpurun_out/r02_sessionP.log 2>&1
cut -c1-900 gpurun_out/r02_sessionP.log | tail -30
