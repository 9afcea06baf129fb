"""N > 1 on real GPUs (needs >= 2 devices; skipped otherwise): torchrun launches
tests/mgpu_worker.py — train_distributed (byte-range shards, NCCL allreduce of the code point
histogram, all-gather of unique words) == oracle, encode_sharded == oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_two_ranks(product, checkers):
    n = product.yttm_device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs (have %d)" % n)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29531", os.path.join(HERE, "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MGPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
