#!/bin/bash
# Round-2 session D (gpurun --gpus N): distributed training on real GPUs (peer stores over NVLink) + bench at N.
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
{
echo "### nvidia-smi topo"; nvidia-smi topo -m 2>&1 | head -12
echo "### mgpu test (train_distributed == oracle, encode_sharded == oracle)"
YTTM_XQ_TIMEOUT_MS=20000 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_worker.py 2>&1 | grep -v "^id:\|^number of\|^model saved\|^Training\|^  \|^$" | tail -25
echo "### mgpu test, YTTM_DRAIN_PLACES=1 (the 8-GPU geometry of the drain)"
YTTM_DRAIN_PLACES=1 YTTM_XQ_TIMEOUT_MS=20000 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29543 tests/mgpu_worker.py 2>&1 | grep "us/merge\|MGPU_OK\|rror" | tail -6
echo "### bench --gpus $N"
YTTM_XQ_TIMEOUT_MS=20000 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02m_bench_n$N.json 2> gpurun_out/r02m_bench_n$N.err; echo "bench rc=$?"
grep -v "^id:\|^number of\|^model saved\|^Training\|^  \|^$" gpurun_out/r02m_bench_n$N.err | tail -15
echo "### probe_train 1 GPU (sweeps per iteration)"; timeout 300 python tools/probe_train.py zipf 32000 100e6 2>&1 | tail -1
} > gpurun_out/r02_sessionM_n$N.log 2>&1
tail -70 gpurun_out/r02_sessionM_n$N.log
