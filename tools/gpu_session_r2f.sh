#!/bin/bash
# Round-2 session F (1 GPU, short): merge loop phase split per block (YTTM_DBG=16) after the revert to the top-1 cache.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "### probe_train 100 MB zipf"; timeout 300 python tools/probe_train.py zipf 32000 100e6 2>&1 | tail -1
echo "### probe_train 100 MB zipf YTTM_DBG=16"; YTTM_DBG=16 timeout 300 python tools/probe_train.py zipf 32000 100e6 2>&1 | grep -E "DBG16|us_per_merge" | tail -14
echo "### probe_train readme YTTM_DBG=16"; YTTM_DBG=16 timeout 300 python tools/probe_train.py readme 2>&1 | grep -E "DBG16|us_per_merge" | tail -14
echo "### ab_encode (dedup with table-resident prefixes)"; timeout 300 python tools/ab_encode.py 1000000 5 gpurun_out/r02f_ab_encode.json 2>&1 | tail -2
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_ab_encode.json'))
for k in ('default','slots','plain'): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in d[k].items()})
PY
} > gpurun_out/r02_sessionF.log 2>&1
tail -60 gpurun_out/r02_sessionF.log
