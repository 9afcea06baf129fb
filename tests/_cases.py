"""Shared parity cases (seeded inputs small enough for the oracle to finish in seconds)."""
import numpy as np

from youtokentome_b200 import synth

EDGE_SENTENCES = [
    b"", b" ", b"   \t\n ", b"a", b" a ", b"a\xff\xfeb \xe2\x96\x81 c\xe2\x96", b"\x80\x80 \xbf", "яы a".encode(),
    b"aaaaaaaaaaaaaaaaaaaa aaaaaaa", b"abcabcabc" * 40, ("▁".join(["ab", "cd", "e"])).encode(),
    b"\xf0\x9f\x98\x80 \xf0\x9f\x98 \xed\xa0\x80 \xc0\xaf z", b"d d", b"a  b   c    d", "日本語 テキスト abc".encode(),
    # words that START with invalid bytes (the BPE-dropout key uses the word's raw start) - found by tools/fuzz_emul.py
    b"\xc0\xafabcabc ab \xff\xfeabab\xff ba \x80b\xe2\x96ab",
]


def stress_case(seed):
    """(train text, vocab_size, coverage, test sentences) in the shape of stress_test.cpp:391-493."""
    rng = np.random.default_rng(seed + 5)
    text = synth.stress_text(seed)
    vocab = len(set(text) | {32}) + 4 + int(rng.integers(0, 40))
    cov = 1.0 if rng.integers(0, 2) == 0 else 1 - rng.random() * 0.4
    sents = [synth.stress_text(seed * 7 + k, 1000 if k == 0 else 30, train=False) for k in range(8)]
    return text, vocab, cov, sents


_zipf = {}


def zipf(n_words=3000, seed=3):
    key = (n_words, seed)
    if key not in _zipf:
        _zipf[key] = synth.ZipfCorpus(n_words=n_words, seed=seed)
    return _zipf[key]


def dirty_zipf_text(n_bytes=200_000):
    """Zipf multi-script text with invalid UTF-8, U+2581 separators and truncated sequences."""
    t = zipf().text(n_bytes)
    h = len(t) // 4
    return t[:h] + b"\xff\xfe \xe2\x96\x81 \xf0\x9f\x98 \xed\xa0\x80 \xc0\xaf" + t[h:]


def zipf_sentences(n=300, target=100, seed=9):
    zc = zipf()
    return zc.sentences(n, target, seed=seed) + [s + b"\xf0\x9f\x98\x80zz" for s in zc.sentences(20, 50, seed=seed + 1)]
