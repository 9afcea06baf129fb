"""Synthetic corpora for the BASELINE.json configs.  Pure numpy / stdlib;
deterministic for a given seed.  Used by tests/ and bench.py (data: "synthetic")."""
import random

import numpy as np

# ---------------------------------------------------------------------------------------------
# config 1: README example / tests/unit_tests/utils_for_testing.py:23-37 of the reference
# ---------------------------------------------------------------------------------------------


def readme_corpus(n_lines=10000, n_chars=100, alphabet="abcd ", seed=19):
    """10 000 lines x 100 chars drawn from "abcd " with random.seed(19) -> bytes (1 010 000 B)."""
    rnd = random.Random(seed)
    lines = ["".join([rnd.choice(alphabet) for _ in range(n_chars)]) for _ in range(n_lines)]
    return ("\n".join(lines) + "\n").encode()


def stress_text(seed, n_limit=1000, train=True):
    """Text with the shape of the reference stress generator (stress_test.cpp:272-311): random
    chars over "abc " (train) / "abcd " (test), no trailing spaces.  Own RNG (numpy)."""
    rng = np.random.default_rng(seed)
    sigma = "abc " if train else "abcd "
    n = int(min(rng.integers(1, 1001), n_limit))
    s = [sigma[0]]
    while len(s) < n:
        if rng.integers(0, 2):
            s.append(sigma[int(rng.integers(0, len(sigma)))])
        else:
            for _ in range(int(rng.integers(1, 5))):
                s.append(sigma[int(rng.integers(0, len(sigma)))])
    s = "".join(s[:n]).rstrip(" ")
    s = s + sigma[0] * (n - len(s))
    return s.encode()


# ---------------------------------------------------------------------------------------------
# configs 2-5: Zipf corpora over a fixed lexicon with a skewed multi-script alphabet
# ---------------------------------------------------------------------------------------------

_SCRIPTS = {
    # name: (first code point, count)
    "latin": (ord("a"), 26),
    "digits": (ord("0"), 10),
    "latin1": (0xE0, 24),        # 2-byte
    "cyrillic": (0x430, 32),     # 2-byte
    "greek": (0x3B1, 24),        # 2-byte
    "cjk": (0x4E00, 3000),       # 3-byte, large alphabet
    "cjk8k": (0x4E00, 8000),     # 3-byte, the ~8k alphabet of BASELINE config 5
    "kana": (0x3041, 80),        # 3-byte
    "deva": (0x905, 50),         # 3-byte
    "arabic": (0x627, 36),       # 2-byte
    "emoji": (0x1F600, 64),      # 4-byte
}


def _script_chars(name):
    a, n = _SCRIPTS[name]
    return np.arange(a, a + n, dtype=np.uint32)


def make_lexicon(n_words=200_000, seed=1234, mix=None, max_len=14):
    """Fixed lexicon: word i = U[1,max_len] chars from one script (chosen by `mix`), chars
    Zipf-skewed inside the script.  Returns list[bytes] (UTF-8)."""
    rng = np.random.default_rng(seed)
    mix = mix or {"latin": 0.70, "digits": 0.03, "latin1": 0.05, "cyrillic": 0.15, "greek": 0.03, "cjk": 0.03,
                  "emoji": 0.01}
    names = list(mix.keys())
    probs = np.array([mix[k] for k in names], dtype=np.float64)
    probs /= probs.sum()
    script_of = rng.choice(len(names), size=n_words, p=probs)
    lens = rng.integers(1, max_len + 1, size=n_words)
    words = []
    seen = set()
    for i in range(n_words):
        chars = _script_chars(names[script_of[i]])
        w = 1.0 / np.arange(1, len(chars) + 1) ** 0.9
        w /= w.sum()
        for _ in range(8):
            cps = rng.choice(chars, size=int(lens[i]), p=w)
            s = "".join(map(chr, cps.tolist())).encode()
            if s not in seen:
                break
            lens[i] += 1
        seen.add(s)
        words.append(s)
    return words


class ZipfCorpus:
    """Words sampled Zipf(s) from a fixed lexicon (configs 2, 3, 4)."""

    def __init__(self, n_words=200_000, s=1.07, seed=1234, mix=None):
        self.lex = make_lexicon(n_words, seed, mix)
        w = 1.0 / np.arange(1, n_words + 1, dtype=np.float64) ** s
        self.cdf = np.cumsum(w / w.sum())
        self.lens = np.array([len(x) for x in self.lex], dtype=np.int64)
        self.seed = seed

    def sample_words(self, rng, n):
        return np.searchsorted(self.cdf, rng.random(n), side="right").clip(0, len(self.lex) - 1)

    def text(self, n_bytes, words_per_line=20, seed=None):
        """~n_bytes of text, `words_per_line` words per '\n'-terminated line -> bytes."""
        rng = np.random.default_rng(self.seed if seed is None else seed)
        avg = float((self.lens * np.diff(np.concatenate([[0.0], self.cdf]))).sum()) + 1.0
        out, total = [], 0
        while total < n_bytes:
            k = int(min(max((n_bytes - total) / avg * 1.02 + 64, 1024), 4_000_000))
            idx = self.sample_words(rng, k)
            parts = []
            for j in range(0, k, words_per_line):
                parts.append(b" ".join(self.lex[i] for i in idx[j:j + words_per_line]))
            chunk = b"\n".join(parts) + b"\n"
            out.append(chunk)
            total += len(chunk)
        return b"".join(out)[:n_bytes].rsplit(b" ", 1)[0] + b"\n"

    def sentences(self, n, target_len=128, seed=None, lognormal=False, clip=(16, 4096)):
        """n sentences of about target_len bytes (cut at a word boundary) -> list[bytes]."""
        rng = np.random.default_rng((self.seed + 1) if seed is None else seed)
        if lognormal:
            sigma = 0.6
            mu = np.log(target_len) - sigma * sigma / 2
            tl = np.clip(rng.lognormal(mu, sigma, n), clip[0], clip[1]).astype(np.int64)
        else:
            tl = np.full(n, target_len, dtype=np.int64)
        out = []
        pool = self.sample_words(rng, 1 << 20)
        pp = 0
        for i in range(n):
            cur, size = [], 0
            while True:
                if pp >= len(pool):
                    pool = self.sample_words(rng, 1 << 20)
                    pp = 0
                w = self.lex[pool[pp]]
                if size + len(w) + (1 if cur else 0) > tl[i]:
                    break
                pp += 1
                cur.append(w)
                size += len(w) + (1 if len(cur) > 1 else 0)
            if not cur:
                cur = [self.lex[pool[pp]][: int(tl[i])]]
                pp += 1
            out.append(b" ".join(cur))
        return out


class FastZipf:
    """Bench-scale generator (vectorised): same idea as ZipfCorpus — Zipf(s) words from a fixed
    lexicon over a skewed multi-script alphabet — but built with numpy only, so 100 MB of text
    or 1 M packed sentences take seconds."""

    def __init__(self, n_words=200_000, s=1.07, seed=1234, mix=None, max_len=14):
        rng = np.random.default_rng(seed)
        mix = mix or {"latin": 0.70, "digits": 0.03, "latin1": 0.05, "cyrillic": 0.15, "greek": 0.03, "cjk": 0.03,
                      "emoji": 0.01}
        names = list(mix.keys())
        probs = np.array([mix[k] for k in names], dtype=np.float64)
        probs /= probs.sum()
        script_of = rng.choice(len(names), size=n_words, p=probs)
        lens = rng.integers(1, max_len + 1, size=n_words)
        words = [None] * n_words
        for si, name in enumerate(names):
            sel = np.nonzero(script_of == si)[0]
            if len(sel) == 0:
                continue
            chars = _script_chars(name)
            w = 1.0 / np.arange(1, len(chars) + 1) ** 0.9
            mat = rng.choice(chars, size=(len(sel), max_len + 6), p=w / w.sum())
            for row, wi in enumerate(sel):
                words[wi] = mat[row]
        seen = set()
        lex = []
        for wi in range(n_words):
            k = int(lens[wi])
            while True:
                key = words[wi][:k].tobytes()
                if key not in seen or k >= max_len + 6:
                    break
                k += 1
            seen.add(key)
            lex.append("".join(map(chr, words[wi][:k].tolist())).encode())
        self.lex = lex
        pw = 1.0 / np.arange(1, n_words + 1, dtype=np.float64) ** s
        self.cdf = np.cumsum(pw / pw.sum())
        self.lens = np.array([len(x) for x in lex], dtype=np.int64)
        self.avg = float((self.lens * np.diff(np.concatenate([[0.0], self.cdf]))).sum()) + 1.0
        self.seed = seed

    def _words(self, rng, k):
        return np.searchsorted(self.cdf, rng.random(k), side="right").clip(0, len(self.lex) - 1)

    def text(self, n_bytes, seed=None, words_per_line=20):
        """About n_bytes of '\\n'-terminated lines of `words_per_line` words -> bytes."""
        rng = np.random.default_rng(self.seed + 11 if seed is None else seed)
        out, total = [], 0
        lex = self.lex
        while total < n_bytes:
            k = int(min((n_bytes - total) / self.avg + 1024, 8_000_000))
            idx = self._words(rng, k).tolist()
            lines = [b" ".join([lex[i] for i in idx[j:j + words_per_line]]) for j in range(0, k, words_per_line)]
            chunk = b"\n".join(lines) + b"\n"
            out.append(chunk)
            total += len(chunk)
        buf = b"".join(out)
        cut = buf.rfind(b"\n", 0, n_bytes)
        return buf[:cut + 1]

    def packed_sentences(self, n, target_len=128, seed=None, lognormal=False, clip=(16, 4096)):
        """n sentences cut at word boundaries from one Zipf word stream -> (bytes, uint64
        offsets[n+1]).  Sentence i ends at the first word boundary at or after its target length
        (fixed, or lognormal with that mean)."""
        rng = np.random.default_rng(self.seed + 23 if seed is None else seed)
        if lognormal:
            sigma = 0.6
            tl = np.clip(rng.lognormal(np.log(target_len) - sigma * sigma / 2, sigma, n), clip[0], clip[1])
        else:
            tl = np.full(n, float(target_len))
        want = np.cumsum(tl)
        k = int(want[-1] / self.avg * 1.05) + 1024
        idx = self._words(rng, k)
        ends = np.cumsum(self.lens[idx] + 1)       # byte position after each word's trailing space
        while ends[-1] < want[-1]:
            more = self._words(rng, k // 10 + 1024)
            ends = np.concatenate([ends, ends[-1] + np.cumsum(self.lens[more] + 1)])
            idx = np.concatenate([idx, more])
        cut = np.searchsorted(ends, want, side="left")        # word index closing each sentence
        cut = np.maximum.accumulate(np.maximum(cut, np.arange(n)))
        n_used = int(cut[-1]) + 1
        lex = self.lex
        buf = b" ".join([lex[i] for i in idx[:n_used].tolist()]) + b" "
        offs = np.zeros(n + 1, dtype=np.uint64)
        offs[1:] = ends[cut]
        return buf, offs


# BASELINE config 5: Latin 40 %, Cyrillic 20 %, CJK 25 % (3-byte, ~8k alphabet), Arabic / Devanagari /
# emoji (4-byte) 15 %; per-script Zipf lexicons, seed 777.  Rare CJK characters fall under coverage 0.9999.
CONFIG5_MIX = {"latin": 0.40, "cyrillic": 0.20, "cjk8k": 0.25, "arabic": 0.06, "deva": 0.06, "emoji": 0.03}

_GEN = {  # bench-scale corpora: kind -> (FastZipf arguments, first chunk seed)
    "zipf": (dict(n_words=200_000, s=1.07, seed=1234), 5000),
    "multilingual": (dict(n_words=300_000, s=1.05, seed=777, mix=CONFIG5_MIX), 7000),
}
_GEN_CACHE = {}


def _gen_chunk(job):
    kind, k, n_bytes, path = job
    import os
    if path and os.path.exists(path) and os.path.getsize(path) > 0:
        with open(path, "rb") as f:
            return f.read()
    args, seed0 = _GEN[kind]
    if kind not in _GEN_CACHE:
        _GEN_CACHE[kind] = FastZipf(**args)
    data = _GEN_CACHE[kind].text(n_bytes, seed=seed0 + k)
    if path:
        tmp = "%s.tmp%d" % (path, os.getpid())
        with open(tmp, "wb") as f:
            f.write(data)
        os.replace(tmp, path)
    return data


def corpus_chunks(kind, chunk_ids, chunk_bytes=125_000_000, cache_dir=None, workers=None):
    """A corpus defined as independently seeded chunks (chunk k: seed0 + k, ~chunk_bytes, ends with a newline), so that
    a rank of a multi-GPU job can make exactly its part and every world size sees the same corpus.  Chunks are
    generated by a process pool (FastZipf.text is single-threaded: ~3 s per 100 MB).  Returns list[bytes]."""
    import os
    from concurrent.futures import ProcessPoolExecutor
    chunk_ids = list(chunk_ids)
    jobs = [(kind, k, chunk_bytes, os.path.join(cache_dir, "%s_%d_%d.bin" % (kind, chunk_bytes, k)) if cache_dir else None)
            for k in chunk_ids]
    workers = workers or max(1, min(len(jobs), (os.cpu_count() or 2) // 2, 16))
    if workers == 1 or len(jobs) == 1:
        return [_gen_chunk(j) for j in jobs]
    with ProcessPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(_gen_chunk, jobs))


MULTILINGUAL_MIX = {"latin": 0.40, "cyrillic": 0.20, "cjk": 0.20, "kana": 0.05, "arabic": 0.05, "deva": 0.05,
                    "emoji": 0.03, "latin1": 0.02}


def multilingual_corpus(n_words=300_000, seed=777):
    """Config 5 shape: Latin 40 %, Cyrillic 20 %, CJK/kana 25 %, Arabic/Devanagari/emoji 15 %."""
    return ZipfCorpus(n_words=n_words, s=1.05, seed=seed, mix=MULTILINGUAL_MIX)


# golden corpora of the reference's tests/unit_tests/test_manual.py (train text, test text, vocab)
GOLDEN_TEXTS = {
    "russian": ("\n        собирать cборник сборище отобранный сборщица \n        ",
                "\n        собранный собрание прибор\n        ", 50),
    "english": ("\n        anachronism\n        synchronous  \n        chronology\n        chronic\n"
                "        chronophilia\n        chronoecological\n        chronocoulometry\n        ",
                "chronocline synchroscope ", 200),
    "japanese": ("\n        むかし、 むかし、 ある ところ に\n        おじいさん と おばあさん が いました。\n"
                 "        おじいさん が 山（やま） へ 木（き） を きり に いけば、\n"
                 "        おばあさん は 川（かわ） へ せんたく に でかけます。\n"
                 "        「おじいさん、 はよう もどって きなされ。」\n        「おばあさん も き を つけて な。」\n"
                 "        まい日（にち） やさしく いい あって でかけます \n    ",
                 " おばあさん が  川 で せん ", 100),
}
