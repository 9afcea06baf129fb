#!/bin/bash
# Round-2 session O (1 GPU, short): A/B of the chunk size of the host-buffer encode pipeline (the e2e headline).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for cfg in "32 32" "16 16" "16 8" "8 8" "8 4" "64 16" "24 8"; do
  set -- $cfg
  echo "### YTTM_ENC_CHUNK_MB=$1 YTTM_ENC_FIRST_CHUNK_MB=$2"
  YTTM_ENC_CHUNK_MB=$1 YTTM_ENC_FIRST_CHUNK_MB=$2 timeout 300 python bench.py --no-train-legs --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'ms', round(d['e2e']['ms_per_step'], 3), 'pageable', d['e2e'].get('pageable_value'))"
done
} > gpurun_out/r02_sessionO.log 2>&1
cat gpurun_out/r02_sessionO.log
