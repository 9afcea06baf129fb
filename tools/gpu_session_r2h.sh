#!/bin/bash
# Round-2 session H (1 GPU, very short): per-block phase table of the merge loop (YTTM_DBG=16), front size A/B.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "### probe_train 100 MB zipf YTTM_DBG=16"; YTTM_DBG=16 timeout 300 python tools/probe_train.py zipf 32000 100e6 2>&1 | grep -E "DBG16" | cut -c1-400
for t in 6 4 3; do
echo "### probe_train 100 MB zipf YTTM_FRONT_TOP=$t"; YTTM_FRONT_TOP=$t timeout 300 python tools/probe_train.py zipf 32000 100e6 2>&1 | tail -1 | cut -c1-600
done
echo "### YTTM_FRONT_TOP=4 YTTM_DBG=16"; YTTM_FRONT_TOP=4 YTTM_DBG=16 timeout 300 python tools/probe_train.py zipf 32000 100e6 2>&1 | grep -E "DBG16" | cut -c1-400
} > gpurun_out/r02_sessionH.log 2>&1
cat gpurun_out/r02_sessionH.log
