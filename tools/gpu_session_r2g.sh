#!/bin/bash
# Round-2 session G (1 GPU, short): merge-loop probes (us per merge, per-block phase table) + H2D staging A/B + train parity.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "### probe_train 100 MB zipf (default threads)"; timeout 300 python tools/probe_train.py zipf 32000 100e6 2>&1 | tail -1
echo "### probe_train 100 MB zipf YTTM_DBG=16"; YTTM_DBG=16 timeout 300 python tools/probe_train.py zipf 32000 100e6 2>&1 | grep -E "DBG16" | tail -8
echo "### probe_train readme"; timeout 300 python tools/probe_train.py readme 2>&1 | tail -1
for t in 4 8 16; do
echo "### probe_train YTTM_TRAIN_PINNED_H2D=$t"; YTTM_TRAIN_PINNED_H2D=$t timeout 300 python tools/probe_train.py zipf 32000 100e6 2>&1 | tail -1 | grep -o '"front_ms.*'
done
echo "### train parity tests"; timeout 900 python -m pytest tests/test_train_gpu.py tests/test_train_scale_gpu.py -x -q -m gpu 2>&1 | tail -3
} > gpurun_out/r02_sessionG.log 2>&1
tail -40 gpurun_out/r02_sessionG.log | cut -c1-1000
