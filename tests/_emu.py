"""TEST HARNESS ONLY: the product's kernels compiled for the CPU under the fiber SIMT emulator
(tests/emul/simt).  Never imported by the package; the product loads libyttm_b200.so and nothing else."""
import importlib.util
import os

from youtokentome_b200 import _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def emu_lib():
    if "L" not in _cache:
        spec = importlib.util.spec_from_file_location("build_emu", os.path.join(_HERE, "emul", "simt", "build_emu.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        import ctypes as C
        _cache["L"] = _lib.bind(C.CDLL(mod.build()))
    return _cache["L"]
