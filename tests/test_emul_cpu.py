"""The device building blocks of youtokentome_b200/csrc/bpe_core.cuh (the same source the
kernels compile) run sequentially on the CPU by tests/emul/emul.cpp and are compared with the
oracle: unit/word detection on raw bytes, run-rule pair counting, arg-max order, in-place rewrite,
per-word encode incl. the BPE-dropout queue model.  A test harness — never a product path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import _cases
from _bind import ROOT, _pack, read_model, tmp_model_path
from youtokentome_b200 import synth


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(ROOT, "tests", "emul", "emul.cpp")
    out = os.path.join(ROOT, "tests", "emul", "libemul.so")
    if not os.path.exists(out) or os.path.getmtime(out) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(ROOT, "youtokentome_b200", "csrc", "bpe_core.cuh"))):
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-x", "c++", "-o", out, src],
                       check=True)
    E = C.CDLL(out)
    E.emul_train.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                             C.c_char_p, C.c_void_p]
    E.emul_encode.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_double,
                              C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    return E


def _emul_encode(E, model, sents, bos=False, eos=False, reverse=False, dropout=0.0, seed=0, first_index=0):
    buf, offs = _pack(sents)
    n = len(sents)
    ids = np.zeros(len(buf) + 3 * n + 8, dtype=np.int32)
    oo = np.zeros(n + 1, dtype=np.uint64)
    assert E.emul_encode(model.encode(), buf, offs.ctypes.data, n, int(bos), int(eos), int(reverse), dropout, seed,
                         first_index, ids.ctypes.data, oo.ctypes.data) == 0
    oo = oo.astype(np.int64)
    return [ids[oo[i]:oo[i + 1]].tolist() for i in range(n)]


def _check(E, oracle, text, vocab, cov, sents):
    m_o, m_e = tmp_model_path("orc"), tmp_model_path("emul")
    try:
        oracle.train(text, m_o, vocab, cov)
    except ValueError:
        assert E.emul_train(text, len(text), vocab, cov, 0, 1, 2, 3, m_e.encode(), None) == 1
        return
    st = np.zeros(4, dtype=np.uint64)
    assert E.emul_train(text, len(text), vocab, cov, 0, 1, 2, 3, m_e.encode(), st.ctypes.data) == 0
    assert read_model(m_o) == read_model(m_e)
    assert int(st[0]) == oracle.last_stats["data_len"]
    enc = oracle.encoder(m_o)
    for kw in [dict(), dict(bos=True, eos=True, reverse=True), dict(dropout=0.3, seed=7, first_index=5),
               dict(dropout=1.0, seed=1)]:
        assert enc.encode(sents, **kw) == _emul_encode(E, m_o, sents, **kw)


@pytest.mark.parametrize("seed", range(40))
def test_stress(emul, oracle, seed):
    text, vocab, cov, sents = _cases.stress_case(seed)
    _check(emul, oracle, text, vocab, cov, sents + _cases.EDGE_SENTENCES)


@pytest.mark.parametrize("cov,vocab", [(1.0, 1500), (0.98, 1500), (0.9, 900)])
def test_dirty_unicode(emul, oracle, cov, vocab):
    _check(emul, oracle, _cases.dirty_zipf_text(), vocab, cov, _cases.zipf_sentences(200))


def test_runs_and_corpora(emul, oracle):
    _check(emul, oracle, b"a" * 500 + b" " + b"ab" * 300 + b" aaa aaaa aaaaa", 30, 1.0, [b"a" * 301, b"ab" * 100 + b"a"])
    for name, (tr, te, vocab) in synth.GOLDEN_TEXTS.items():
        _check(emul, oracle, tr.encode(), vocab, 1.0, [te.encode()])
