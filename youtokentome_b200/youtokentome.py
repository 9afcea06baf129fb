"""The reference's Python surface (youtokentome/youtokentome.py:1-99 + the Cython class
youtokentome/cpp/yttm.pyx:52-181) over the B200 library: same class, method names, argument
meaning, return types and exceptions (ValueError(status.message), TypeError for bad argument
types).  Additions are additive only: `encode_packed` (zero-marshalling numpy path) and
`dropout_seed`."""
import ctypes as C
from collections.abc import Collection
from enum import Enum
from typing import List, Optional, Union

import numpy as np

from . import _lib


class OutputType(Enum):
    ID = 1
    SUBWORD = 2


def _pack(sentences):
    enc = [s.encode() if isinstance(s, str) else bytes(s) for s in sentences]
    offs = np.zeros(len(enc) + 1, dtype=np.uint64)
    if enc:
        np.cumsum([len(b) for b in enc], out=offs[1:])
    return b"".join(enc), offs


class BPE:
    def __init__(self, model: str, n_threads: int = -1):
        self.model = model
        self.n_threads = n_threads
        self._open()

    def _open(self):
        L = _lib.lib()
        self._h = L.yttm_api_open(self.model.encode(), self.n_threads)
        if not self._h:
            raise ValueError(L.yttm_api_last_error(None).decode())

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().yttm_api_close(h)
            except Exception:
                pass
            self._h = None

    def _err(self):
        return ValueError(_lib.lib().yttm_api_last_error(self._h).decode())

    @staticmethod
    def train(data: str, model: str, vocab_size: int, coverage: float = 1.0, n_threads: int = -1, pad_id: int = 0,
              unk_id: int = 1, bos_id: int = 2, eos_id: int = 3) -> "BPE":
        L = _lib.lib()
        rc = L.yttm_api_train(data.encode(), model.encode(), vocab_size, coverage, n_threads, pad_id, unk_id, bos_id,
                              eos_id)
        if rc != 0:
            raise ValueError(L.yttm_api_last_error(None).decode())
        return BPE(model=model, n_threads=n_threads)

    # -- encode ---------------------------------------------------------------------------------
    def encode_packed(self, data: bytes, offsets: np.ndarray, bos=False, eos=False, reverse=False, dropout_prob=0.0):
        """Additive fast path: sentence i = data[offsets[i]:offsets[i+1]] (uint64 offsets).
        Returns (int32 ids, uint64 id_offsets) as numpy arrays."""
        L = _lib.lib()
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        total = C.c_uint64(0)
        if isinstance(data, np.ndarray):
            ptr = np.ascontiguousarray(data).ctypes.data
        else:
            data = bytes(data) if not isinstance(data, bytes) else data
            ptr = C.cast(C.c_char_p(data), C.c_void_p)
        rc = L.yttm_api_encode_ids(self._h, ptr, offsets.ctypes.data, n, int(bos), int(eos), int(reverse),
                                   float(dropout_prob), C.byref(total))
        if rc != 0:
            raise self._err()
        ids = np.empty(max(total.value, 1), dtype=np.int32)
        oo = np.empty(n + 1, dtype=np.uint64)
        L.yttm_api_result_ids(self._h, ids.ctypes.data, oo.ctypes.data)
        return ids[:total.value], oo

    def encode(self, sentences: Union[str, List[str]], output_type: OutputType = OutputType.ID, bos: bool = False,
               eos: bool = False, reverse: bool = False, dropout_prob: float = 0):
        if not isinstance(output_type, OutputType):
            raise TypeError("parameter output_type must be youtokentome.OutputType, not %s}" % str(type(output_type)))
        if dropout_prob < 0 or dropout_prob > 1:  # yttm.pyx:92-93
            raise ValueError("dropout_prob value must be in the range [0, 1]. Current value of dropout_prob = " +
                             str(dropout_prob))
        single = isinstance(sentences, str)
        if not single:
            assert isinstance(sentences, (list, tuple))
        data, offs = _pack([sentences] if single else sentences)
        L = _lib.lib()
        if output_type == OutputType.ID:
            ids, oo = self.encode_packed(data, offs, bos, eos, reverse, dropout_prob)
            oo = oo.astype(np.int64)
            flat = ids.tolist()
            out = [flat[oo[i]:oo[i + 1]] for i in range(len(oo) - 1)]
        else:
            need = L.yttm_api_encode_subwords(self._h, data, offs.ctypes.data, len(offs) - 1, int(bos), int(eos),
                                              int(reverse), float(dropout_prob))
            if need < 0:
                raise self._err()
            buf = C.create_string_buffer(int(need) + 1)
            L.yttm_api_result_text(self._h, buf)
            text = buf.raw[:need].decode()
            out = [ln.split("\x01") if ln else [] for ln in text.split("\n")[:-1]]
        return out[0] if single else out

    # -- tables ---------------------------------------------------------------------------------
    def vocab_size(self) -> int:
        return _lib.lib().yttm_api_vocab_size(self._h)

    def vocab(self) -> List[str]:
        L = _lib.lib()
        need = L.yttm_api_vocab(self._h)
        buf = C.create_string_buffer(int(need) + 1)
        L.yttm_api_result_text(self._h, buf)
        return buf.raw[:need].decode().split("\x01")

    def subword_to_id(self, subword: str) -> int:
        return _lib.lib().yttm_api_subword_to_id(self._h, subword.encode())

    def id_to_subword(self, id: int) -> str:
        L = _lib.lib()
        need = L.yttm_api_id_to_subword(self._h, id)
        if need < 0:
            raise self._err()
        buf = C.create_string_buffer(int(need) + 1)
        L.yttm_api_result_text(self._h, buf)
        return buf.raw[:need].decode()

    def decode(self, ids: Union[List[int], List[List[int]]], ignore_ids: Optional[Collection] = None) -> List[str]:
        if not isinstance(ids, list):  # yttm.pyx:138-146
            raise TypeError("{} is not a list instance".format(type(ids)))
        if not isinstance(ignore_ids, Collection) and ignore_ids is not None:
            raise TypeError("{} is not a Collection instance".format(type(ignore_ids)))
        if len(ids) > 0 and isinstance(ids[0], int):
            ids = [ids]
        ign = np.asarray(sorted(ignore_ids) if ignore_ids else [], dtype=np.int32)
        offs = np.zeros(len(ids) + 1, dtype=np.uint64)
        if ids:
            np.cumsum([len(s) for s in ids], out=offs[1:])
        flat = np.asarray([t for s in ids for t in s], dtype=np.int32)
        L = _lib.lib()
        need = L.yttm_api_decode(self._h, flat.ctypes.data, offs.ctypes.data, len(ids), ign.ctypes.data, len(ign))
        if need < 0:
            raise self._err()
        buf = C.create_string_buffer(int(need) + 1)
        L.yttm_api_result_text(self._h, buf)
        # sentences may contain '\n' only if a piece does; pieces never hold whitespace
        return buf.raw[:need].decode().split("\n")[:-1]

    # -- BPE-dropout stream ---------------------------------------------------------------------
    def dropout_seed(self, seed: int):
        """Reset the counter-based dropout generator (seed, sentence counter := 0)."""
        _lib.lib().yttm_api_set_dropout_seed(self._h, seed)

    # -- CLI helpers used by yttm_cli ------------------------------------------------------------
    def encode_cli(self, output_type, stream, bos, eos, reverse, dropout_prob):
        if _lib.lib().yttm_api_encode_cli(self._h, output_type.encode(), int(stream), int(bos), int(eos), int(reverse),
                                          float(dropout_prob)) != 0:
            raise self._err()

    def decode_cli(self, ignore_ids):
        ign = np.asarray(sorted(ignore_ids) if ignore_ids else [], dtype=np.int32)
        if _lib.lib().yttm_api_decode_cli(self._h, ign.ctypes.data, len(ign)) != 0:
            raise self._err()

    def vocab_cli(self, verbose):
        _lib.lib().yttm_api_vocab_cli(self._h, int(verbose))

    # -- pickling (youtokentome.py:90-99) ---------------------------------------------------------
    def __getstate__(self):
        return {"model": self.model, "n_threads": self.n_threads}

    def __setstate__(self, d):
        self.model = d["model"]
        self.n_threads = d["n_threads"]
        self._open()


def release_training_cache():
    """BPE.train keeps its device buffers (corpus, word table, packed words, pair table) cached per host thread so
    that repeated trainings do not reallocate; this gives the calling thread's back to the device."""
    _lib.lib().yttm_api_release_training_cache()


def train_report():
    """Sizes / stage timings of the last BPE.train on this thread (dict)."""
    out = (C.c_double * 16)()
    n = _lib.lib().yttm_api_train_report(out, 16)
    names = ["n_bytes", "data_len", "n_words", "n_unique", "n_tokens", "n_pairs", "n_merges", "read_s", "h2d_ms",
             "char_hist_ms", "word_count_ms", "tokenise_ms", "pair_hist_ms", "merge_loop_ms", "total_s", "launches"]
    return dict(zip(names[:n], list(out)[:n]))
