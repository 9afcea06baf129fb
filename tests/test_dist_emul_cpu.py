"""Multi-GPU training on the CPU SIMT emulator: `world` ranks = `world` host threads of this process, each driving its
own context of the emulated library through youtokentome_b200.distributed.train_distributed (the product's own host
code; only the transport of the few collectives is replaced by tests/_threadcomm.py).  The merge loops of the ranks
run concurrently and exchange the count changes of every merge through each other's exchange buffers, exactly as the
GPUs do over NVLink.  The model must equal the oracle's (= the reference's) for every world size."""
import os

import pytest

import _bind
import _cases
from _bind import read_model, tmp_model_path
from _threadcomm import run_ranks


@pytest.fixture(scope="module")
def emu():
    from _emu import emu_lib
    return emu_lib()


def _train(emu, world, text, vocab, cov=1.0, specials=(0, 1, 2, 3)):
    from youtokentome_b200 import distributed as D
    model = tmp_model_path("dist")
    pad, unk, bos, eos = specials

    def rank_fn(comm):
        st = {}
        n = D.train_distributed(text, model, vocab, cov, pad, unk, bos, eos, comm=comm, device=0, lib=emu, stats_out=st)
        return n, st

    res = run_ranks(world, rank_fn)
    assert len({r[0] for r in res}) == 1, "ranks disagree on the number of merges"
    return model, res


def _oracle_model(oracle, text, vocab, cov=1.0, specials=(0, 1, 2, 3)):
    m = tmp_model_path("orc")
    pad, unk, bos, eos = specials
    oracle.train(text, m, vocab, cov, pad=pad, unk=unk, bos=bos, eos=eos)
    return m


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_training_equals_the_oracle(emu, oracle, monkeypatch, world):
    monkeypatch.setenv("YT_EMU_SMS", "2")
    monkeypatch.setenv("YTTM_XQ_TIMEOUT_MS", "120000")
    for text, vocab, cov in [(_cases.zipf().text(30_000), 600, 1.0), (_cases.dirty_zipf_text(30_000), 400, 0.98)]:
        m, res = _train(emu, world, text, vocab, cov)
        assert read_model(m) == read_model(_oracle_model(oracle, text, vocab, cov))
        # equal words of different shards met on one rank: the ranks' unique words are disjoint
        single, _ = _train(emu, 1, text, vocab, cov)
        assert read_model(single) == read_model(m)


@pytest.mark.parametrize("sms,places", [("1", "1"), ("2", None), ("3", "2")])
def test_distributed_training_eight_ranks(emu, oracle, monkeypatch, sms, places):
    """The largest world (XQ_MAX_WORLD = 8): one block per rank, the drain geometry of an 8-GPU job (one place per
    segment handled by the items, the rest by the shared walk), one count word per remote rank."""
    monkeypatch.setenv("YT_EMU_SMS", sms)
    if places:
        monkeypatch.setenv("YTTM_DRAIN_PLACES", places)
    monkeypatch.setenv("YTTM_XQ_TIMEOUT_MS", "240000")
    for text, vocab, cov in [(_cases.zipf().text(20_000), 400, 1.0), (_cases.dirty_zipf_text(20_000), 300, 0.98)]:
        m, res = _train(emu, 8, text, vocab, cov)
        assert read_model(m) == read_model(_oracle_model(oracle, text, vocab, cov))


def test_distributed_words_are_deduplicated_across_ranks(emu, oracle, monkeypatch):
    """Two ranks whose shards hold the SAME words: after the exchange the job holds each word once (the round-1 path
    kept one copy per rank, which made the replicated loop N times larger)."""
    monkeypatch.setenv("YT_EMU_SMS", "2")
    monkeypatch.setenv("YTTM_XQ_TIMEOUT_MS", "120000")
    half = _cases.zipf().text(12_000)
    text = half + b" " + half
    _, res1 = _train(emu, 1, text, 300)
    _, res2 = _train(emu, 2, text, 300)
    assert sum(r[1]["n_unique"] for r in res2) == res1[0][1]["n_unique"]
    assert sum(r[1]["n_tokens"] for r in res2) == res1[0][1]["n_tokens"]


def test_distributed_rare_paths(emu, oracle, monkeypatch):
    """Tiny exchange segments (overflow -> every rank leaves for a rebuild at the same merge), tiny table partitions
    (chunked table exchange, growth in lockstep), forced STREAMING tiles, a shard without words."""
    monkeypatch.setenv("YT_EMU_SMS", "2")
    monkeypatch.setenv("YTTM_XQ_TIMEOUT_MS", "120000")
    monkeypatch.setenv("YTTM_XQ_SEG_CAP", "8")
    monkeypatch.setenv("YTTM_PAIR_CAP_FLOOR", "16")
    text = _cases.zipf().text(20_000)
    m, _ = _train(emu, 2, text, 500)
    assert read_model(m) == read_model(_oracle_model(oracle, text, 500))
    monkeypatch.setenv("YTTM_FORCE_STREAM", "1")
    monkeypatch.setenv("YTTM_STREAM_Q", "128")
    m, _ = _train(emu, 2, text, 300)
    assert read_model(m) == read_model(_oracle_model(oracle, text, 300))
    monkeypatch.delenv("YTTM_FORCE_STREAM")
    lopsided = b"abab abab cd " + b" " * 4000   # the second byte range holds no word at all
    m, _ = _train(emu, 2, lopsided, 12)
    assert read_model(m) == read_model(_oracle_model(oracle, lopsided, 12))


def test_distributed_errors_reach_every_rank(emu, monkeypatch):
    """check_config messages of the reference, and a failure on every rank at once instead of a hang."""
    monkeypatch.setenv("YT_EMU_SMS", "2")
    from youtokentome_b200 import distributed as D
    with pytest.raises(ValueError, match="coverage value must be in the range"):
        D.check_config(100, 1.5, 0, 1, 2, 3)
    with pytest.raises(ValueError, match="All ids of special tokens must be different"):
        D.check_config(100, 1.0, 0, 1, 1, 3)
    with pytest.raises(ValueError, match="unk_id: must be in the range"):
        D.check_config(100, 1.0, 0, -1, 2, 3)
    with pytest.raises(ValueError, match="Vocabulary size too small"):
        _train(emu, 2, _cases.zipf().text(5_000), 10)
