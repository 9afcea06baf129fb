"""N > 1 on real GPUs (skipped when the box has fewer devices): torchrun launches tests/mgpu_worker.py, one process per
GPU over NCCL — distributed.train_distributed (byte-range shards, NCCL allreduce of the code point histogram, unique
words hash-partitioned across the ranks by an all-to-all of device buffers, table build and per-merge count exchange by
peer stores inside the kernels) == oracle for coverage 1.0 / 0.98, all ranks return the same rules, and
encode_sharded == oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_ranks(product, checkers, world):
    n = product.yttm_device_count()
    if n < world:
        pytest.skip("needs %d GPUs (have %d)" % (world, n))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world, "--master-addr",
           "127.0.0.1", "--master-port", str(29531 + world), os.path.join(HERE, "mgpu_worker.py")]
    env = dict(os.environ, YTTM_XQ_TIMEOUT_MS="20000")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert "MGPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
