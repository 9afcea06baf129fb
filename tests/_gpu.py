"""Helpers that call the PRODUCT (CUDA) path through its C ABI / Python surface."""
import ctypes as C

import numpy as np

from _bind import _pack, _unpack, tmp_model_path
from youtokentome_b200 import _lib


def gpu_train(text, vocab, cov=1.0, pad=0, unk=1, bos=2, eos=3, model=None):
    """learn_bpe_from_string on the GPU -> model path.  Raises ValueError(Status.message)."""
    L = _lib.lib()
    model = model or tmp_model_path("gpu")
    rc = L.yttm_api_train_memory(text, len(text), model.encode(), vocab, cov, pad, unk, bos, eos)
    if rc != 0:
        raise ValueError(L.yttm_api_last_error(None).decode())
    return model


class GpuEncoder:
    def __init__(self, model):
        L = _lib.lib()
        self.h = L.yttm_api_open(model.encode(), 1)
        if not self.h:
            raise ValueError(L.yttm_api_last_error(None).decode())

    def __del__(self):
        try:
            _lib.lib().yttm_api_close(self.h)
        except Exception:
            pass

    def encode(self, sentences, bos=False, eos=False, reverse=False, dropout=0.0, seed=None):
        L = _lib.lib()
        if seed is not None:
            L.yttm_api_set_dropout_seed(self.h, seed)
        buf, offs = _pack(sentences)
        tot = C.c_uint64(0)
        rc = L.yttm_api_encode_ids(self.h, C.cast(C.c_char_p(buf), C.c_void_p), offs.ctypes.data, len(sentences),
                                   int(bos), int(eos), int(reverse), dropout, C.byref(tot))
        if rc != 0:
            raise ValueError(L.yttm_api_last_error(self.h).decode())
        ids = np.zeros(max(tot.value, 1), dtype=np.int32)
        oo = np.zeros(len(sentences) + 1, dtype=np.uint64)
        L.yttm_api_result_ids(self.h, ids.ctypes.data, oo.ctypes.data)
        return _unpack(ids[:tot.value], oo)
