"""ctypes bindings for the CHECKERS (test infrastructure): oracle/liboracle.so (our CPU
restatement) and oracle/_ref/libyttm_ref_{det,prod}.so (the unmodified reference compiled by
oracle/Makefile).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ERRLEN = 1024


def build_checkers():
    """Compile oracle/liboracle.so and, when /root/reference exists, oracle/_ref/*.so."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True, stdout=subprocess.DEVNULL)


def _pack(sentences):
    """list[bytes] -> (concatenated bytes buffer, uint64 offsets[n+1])."""
    offs = np.zeros(len(sentences) + 1, dtype=np.uint64)
    if sentences:
        offs[1:] = np.cumsum([len(s) for s in sentences], dtype=np.uint64)
    return b"".join(sentences), offs


def _unpack(ids, offs):
    offs = offs.astype(np.int64)
    return [ids[offs[i]:offs[i + 1]].tolist() for i in range(len(offs) - 1)]


def read_model(path):
    """Model file (utils.cpp:50-91) -> (char2id dict, rules list[(x,y,z)], (unk,pad,bos,eos))."""
    with open(path) as f:
        tok = f.read().split()
    n, m = int(tok[0]), int(tok[1])
    p = 2
    c2i = {}
    for _ in range(n):
        c2i[int(tok[p])] = int(tok[p + 1])
        p += 2
    rules = []
    for _ in range(m):
        rules.append((int(tok[p]), int(tok[p + 1]), int(tok[p + 2])))
        p += 3
    special = tuple(int(x) for x in tok[p:p + 4])
    return c2i, rules, special


class _Enc:
    """Common encoder wrapper; subclasses set prefix / call conventions."""

    def __init__(self, lib, handle, prefix):
        self.lib, self.h, self.prefix = lib, handle, prefix

    def __del__(self):
        try:
            getattr(self.lib, self.prefix + "_encoder_free")(self.h)
        except Exception:
            pass

    def vocab_size(self):
        return getattr(self.lib, self.prefix + "_vocab_size")(self.h)


class Reference:
    """The unmodified reference (oracle/_ref).  kind: 'det' (parity) or 'prod' (timing)."""

    def __init__(self, kind="det"):
        path = os.path.join(ORACLE_DIR, "_ref", "libyttm_ref_%s.so" % kind)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = lib = C.CDLL(path)
        lib.ref_train_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.POINTER(C.c_double), C.c_char_p, C.c_int]
        lib.ref_train_memory.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_char_p, C.c_int]
        lib.ref_encoder_new.restype = C.c_void_p
        lib.ref_encoder_new.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        lib.ref_encoder_free.argtypes = [C.c_void_p]
        lib.ref_vocab_size.argtypes = [C.c_void_p]
        lib.ref_encode_ids.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                       C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_char_p, C.c_int]
        lib.ref_result_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_encode_subwords.restype = C.c_int64
        lib.ref_encode_subwords.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                            C.c_char_p, C.c_int64]
        lib.ref_decode_ids.restype = C.c_int64
        lib.ref_decode_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_char_p, C.c_int64]
        self.deterministic = bool(lib.ref_is_deterministic_queue())

    def train(self, text, model_path, vocab_size, coverage=1.0, n_threads=1, pad=0, unk=1, bos=2, eos=3):
        """text: bytes.  Returns seconds; raises ValueError(Status.message)."""
        err = C.create_string_buffer(ERRLEN)
        sec = C.c_double(0)
        rc = self.lib.ref_train_memory(text, len(text), model_path.encode(), vocab_size, coverage, n_threads, pad, unk,
                                       bos, eos, C.byref(sec), err, ERRLEN)
        if rc != 0:
            raise ValueError(err.value.decode())
        return sec.value

    def train_file(self, input_path, model_path, vocab_size, coverage=1.0, n_threads=1, pad=0, unk=1, bos=2, eos=3):
        err = C.create_string_buffer(ERRLEN)
        sec = C.c_double(0)
        rc = self.lib.ref_train_file(input_path.encode(), model_path.encode(), vocab_size, coverage, n_threads, pad,
                                     unk, bos, eos, C.byref(sec), err, ERRLEN)
        if rc != 0:
            raise ValueError(err.value.decode())
        return sec.value

    def encoder(self, model_path, n_threads=1):
        err = C.create_string_buffer(ERRLEN)
        h = self.lib.ref_encoder_new(model_path.encode(), n_threads, err, ERRLEN)
        if not h:
            raise ValueError(err.value.decode())
        return RefEncoder(self.lib, h)


class RefEncoder(_Enc):
    def __init__(self, lib, h):
        super().__init__(lib, h, "ref")
        self.last_seconds = 0.0

    def encode_packed(self, buf, offs, bos=False, eos=False, reverse=False, dropout=0.0):
        err = C.create_string_buffer(ERRLEN)
        sec, tot = C.c_double(0), C.c_uint64(0)
        n = len(offs) - 1
        rc = self.lib.ref_encode_ids(self.h, buf, offs.ctypes.data, n, int(bos), int(eos), int(reverse), dropout,
                                     C.byref(sec), C.byref(tot), err, ERRLEN)
        if rc != 0:
            raise ValueError(err.value.decode())
        self.last_seconds = sec.value
        ids = np.zeros(max(tot.value, 1), dtype=np.int32)
        oo = np.zeros(n + 1, dtype=np.uint64)
        self.lib.ref_result_ids(self.h, ids.ctypes.data, oo.ctypes.data)
        return ids[:tot.value], oo

    def encode(self, sentences, **kw):
        buf, offs = _pack(sentences)
        ids, oo = self.encode_packed(buf, offs, **kw)
        return _unpack(ids, oo)

    def encode_subwords(self, sentences, bos=False, eos=False, reverse=False):
        buf, offs = _pack(sentences)
        need = self.lib.ref_encode_subwords(self.h, buf, offs.ctypes.data, len(sentences), int(bos), int(eos),
                                            int(reverse), None, 0)
        out = C.create_string_buffer(int(need) + 1)
        self.lib.ref_encode_subwords(self.h, buf, offs.ctypes.data, len(sentences), int(bos), int(eos), int(reverse),
                                     out, need)
        text = out.raw[:need].decode("utf-8", errors="replace")
        return [ln.split("\x01") if ln else [] for ln in text.split("\n")[:-1]]

    def decode(self, ids):
        a = np.asarray(ids, dtype=np.int32)
        need = self.lib.ref_decode_ids(self.h, a.ctypes.data, len(a), None, 0)
        out = C.create_string_buffer(int(need) + 1)
        self.lib.ref_decode_ids(self.h, a.ctypes.data, len(a), out, need)
        return out.raw[:need].decode("utf-8", errors="replace")


class Oracle:
    """oracle/liboracle.so — our CPU restatement."""

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_checkers()
        self.lib = lib = C.CDLL(path)
        lib.orc_train_memory.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p, C.c_char_p, C.c_int]
        lib.orc_encoder_new.restype = C.c_void_p
        lib.orc_encoder_new.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        lib.orc_encoder_free.argtypes = [C.c_void_p]
        lib.orc_vocab_size.argtypes = [C.c_void_p]
        lib.orc_encode_ids.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                       C.c_double, C.c_uint64, C.c_uint64, C.POINTER(C.c_double),
                                       C.POINTER(C.c_uint64), C.c_char_p, C.c_int]
        lib.orc_result_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.last_stats = None

    def train(self, text, model_path, vocab_size, coverage=1.0, pad=0, unk=1, bos=2, eos=3):
        err = C.create_string_buffer(ERRLEN)
        sec = C.c_double(0)
        stats = np.zeros(5, dtype=np.uint64)
        rc = self.lib.orc_train_memory(text, len(text), model_path.encode(), vocab_size, coverage, pad, unk, bos, eos,
                                       C.byref(sec), stats.ctypes.data, err, ERRLEN)
        if rc != 0:
            raise ValueError(err.value.decode())
        self.last_stats = dict(zip(["data_len", "n_words", "n_unique", "n_tokens", "n_merges"], stats.tolist()))
        return sec.value

    def encoder(self, model_path):
        err = C.create_string_buffer(ERRLEN)
        h = self.lib.orc_encoder_new(model_path.encode(), err, ERRLEN)
        if not h:
            raise ValueError(err.value.decode())
        return OracleEncoder(self.lib, h)


class OracleEncoder(_Enc):
    def __init__(self, lib, h):
        super().__init__(lib, h, "orc")
        self.last_seconds = 0.0

    def encode_packed(self, buf, offs, bos=False, eos=False, reverse=False, dropout=0.0, seed=0, first_index=0):
        err = C.create_string_buffer(ERRLEN)
        sec, tot = C.c_double(0), C.c_uint64(0)
        n = len(offs) - 1
        rc = self.lib.orc_encode_ids(self.h, buf, offs.ctypes.data, n, int(bos), int(eos), int(reverse), dropout, seed,
                                     first_index, C.byref(sec), C.byref(tot), err, ERRLEN)
        if rc != 0:
            raise ValueError(err.value.decode())
        self.last_seconds = sec.value
        ids = np.zeros(max(tot.value, 1), dtype=np.int32)
        oo = np.zeros(n + 1, dtype=np.uint64)
        self.lib.orc_result_ids(self.h, ids.ctypes.data, oo.ctypes.data)
        return ids[:tot.value], oo

    def encode(self, sentences, **kw):
        buf, offs = _pack(sentences)
        ids, oo = self.encode_packed(buf, offs, **kw)
        return _unpack(ids, oo)


def have_reference(kind="det"):
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libyttm_ref_%s.so" % kind))


def tmp_model_path(tag="m"):
    d = tempfile.mkdtemp(prefix="yttm_b200_")
    return os.path.join(d, tag + ".yttm")
