"""torchrun worker (one process per GPU): train_distributed + encode_sharded against the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import _cases  # noqa: E402
from _bind import Oracle, _pack, read_model, tmp_model_path  # noqa: E402
import youtokentome_b200 as yttm  # noqa: E402
from youtokentome_b200 import distributed as D  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ["YTTM_DEVICE"] = str(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    text = _cases.dirty_zipf_text(2_000_000)
    model = "/tmp/yttm_mgpu_%d.yttm" % world
    for cov in (1.0, 0.98):
        st = {}
        n = D.train_distributed(text, model, 3000, coverage=cov, stats_out=st)
        # every rank elected the same merges, and the job's unique words are a partition (each word on one rank)
        t = torch.tensor([n, st["n_unique"], st["n_tokens"]], dtype=torch.int64, device="cuda")
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        assert len({int(x[0]) for x in g}) == 1, "ranks disagree on the number of merges"
        if rank == 0:
            print("unique words per rank:", [int(x[1]) for x in g], "tokens per rank:", [int(x[2]) for x in g],
                  "us/merge %.2f" % (st["merge_loop_ms"] * 1e3 / max(n, 1)), st["phase_us_per_iter"])
        if rank == 0:
            orc = Oracle()
            m = tmp_model_path("orc")
            orc.train(text, m, 3000, cov)
            assert read_model(m) == read_model(model), "distributed training differs from the oracle (cov %s)" % cov
            print("train_distributed ok: world %d, %d merges, coverage %s" % (world, n, cov))
        dist.barrier()
    sents = _cases.zipf_sentences(4000) + _cases.EDGE_SENTENCES
    buf, offs = _pack(sents)
    bpe = yttm.BPE(model)
    lo, hi, ids, oo = D.encode_sharded(bpe, buf, offs, bos=True)
    parts = [None] * world
    dist.all_gather_object(parts, (lo, hi, ids.tolist(), oo.astype(np.int64).tolist()))
    if rank == 0:
        got = []
        for lo_, hi_, ids_, oo_ in parts:
            got += [ids_[oo_[i]:oo_[i + 1]] for i in range(hi_ - lo_)]
        want = Oracle().encoder(model).encode(sents, bos=True)
        assert got == want, "sharded encode differs from the oracle"
        print("encode_sharded ok: %d sentences over %d ranks" % (len(sents), world))
        print("MGPU_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
