"""Probe: train on a synthetic corpus through the device ABI and print the merge-loop phase split."""
import ctypes as C, sys, os, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from youtokentome_b200 import _lib, synth
L = _lib.lib()
which = sys.argv[1] if len(sys.argv) > 1 else "zipf"
vocab = int(sys.argv[2]) if len(sys.argv) > 2 else 32000
if which == "readme":
    text = synth.readme_corpus(); vocab = 5000
elif which == "multilingual":   # a chunk of bench.py's config-5 corpus
    text = b"".join(synth.corpus_chunks("multilingual", [0], int(float(sys.argv[3]) if len(sys.argv) > 3 else 125e6)))
else:
    text = synth.FastZipf().text(int(float(sys.argv[3]) if len(sys.argv) > 3 else 100e6))
ctx = C.c_void_p(); assert L.yttm_ctx_create(0, C.byref(ctx)) == 0
for rep in range(2):
    dl, nd = C.c_uint64(0), C.c_uint64(0)
    assert L.yttm_train_load_corpus(ctx, C.cast(C.c_char_p(text), C.c_void_p), len(text), 0) == 0
    assert L.yttm_train_char_hist(ctx, C.byref(dl), C.byref(nd)) == 0
    cps = np.zeros(nd.value, dtype=np.uint32); cnt = np.zeros(nd.value, dtype=np.uint64)
    L.yttm_train_get_char_hist(ctx, cps.ctypes.data, cnt.ctypes.data)
    order = np.lexsort((cps, cnt))[::-1]
    kc = np.concatenate([[9601], cps[order]]).astype(np.uint32); ki = np.arange(4, 4 + len(kc), dtype=np.uint32)
    assert L.yttm_train_set_alphabet(ctx, kc.ctypes.data, ki.ctypes.data, len(kc), 4) == 0
    st = _lib.TrainStats(); assert L.yttm_train_build(ctx, C.byref(st)) == 0, L.yttm_last_error(ctx)
    nm = vocab - 4 - len(kc); rules = np.zeros(3 * nm, dtype=np.uint32); fr = np.zeros(nm, dtype=np.uint64); nd2 = C.c_uint32(0)
    t0 = time.perf_counter()
    assert L.yttm_train_run(ctx, 4 + len(kc), nm, rules.ctypes.data, fr.ctypes.data, C.byref(nd2)) == 0, L.yttm_last_error(ctx)
    wall = time.perf_counter() - t0
    g = lambda k: L.yttm_stage_ms(ctx, k.encode())
    it = max(g("loop_iters"), 1)
    print(json.dumps({"corpus": which, "bytes": len(text), "U": st.n_unique, "T": st.n_tokens, "P0": st.n_pairs, "cap": g("table_capacity"),
        "merges": nd2.value, "loop_ms": g("merge_loop"), "wall_ms": wall * 1e3, "us_per_merge": g("merge_loop") * 1e3 / max(nd2.value, 1),
        "launches": g("loop_launches"), "refreshes": g("loop_refreshes"), "phase_us_per_iter": {k: g(k) * 1e3 / it for k in ["loop_elect", "loop_apply", "loop_partition", "loop_drain"]},
        "front_ms": {k: g(k) for k in ["h2d", "char_hist", "word_count", "tokenise", "pair_hist"]}}))
L.yttm_ctx_destroy(ctx)
