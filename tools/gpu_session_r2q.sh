#!/bin/bash
# Round-2 session Q (1 GPU, short): the STREAMING recount test incl. the new heavy-merge case.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{ timeout 600 python -m pytest tests/test_train_scale_gpu.py -x -q -m gpu 2>&1 | tail -5; } > gpurun_out/r02_sessionQ.log 2>&1
cat gpurun_out/r02_sessionQ.log
