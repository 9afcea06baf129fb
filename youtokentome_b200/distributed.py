"""One-process-per-GPU plumbing (torch.distributed) around the C ABI — SURVEY.md §8e.

Encode shards by sentence: contiguous ranges balanced by bytes, NO collective on the data path
(`shard_sentences`, `encode_sharded`).

Training shards the corpus by byte range cut at ASCII spaces exactly as the reference cuts it
between threads (bpe.cpp:864-873): every rank runs the two byte passes (code point histogram,
word split + dedup) on its shard on its own GPU; the dense uint64 histogram is sum-allreduced
in place in device memory (NCCL), the unique words of all ranks are all-gathered (duplicates
across ranks are harmless — every statistic is additive in the word frequency, the same reason
the reference can sum per-thread maps, bpe.cpp:1029-1039) and the merge loop then runs
replicated on every rank: at L2/SMEM-resident sizes it is latency-bound (two grid barriers per
merge), so splitting it would only add a per-merge collective.
"""
import ctypes as C

import numpy as np


# ---------------------------------------------------------------------------------------------
# pure host logic (unit-tested on CPU with gloo, world_size 2)
# ---------------------------------------------------------------------------------------------
def _is_space_byte(b):
    return b == 32 or 9 <= b <= 13


def split_byte_ranges(data, world):
    """split_pos of learn_bpe_from_string (bpe.cpp:864-873): boundary i = n*i/world advanced to
    the next ASCII space byte.  Returns world+1 positions."""
    n = len(data)
    pos = [0]
    for i in range(1, world + 1):
        c = n * i // world
        while c < n and not _is_space_byte(data[c]):
            c += 1
        pos.append(c)
    return pos


def shard_sentences(offsets, rank, world):
    """Contiguous sentence range [lo, hi) of `rank`, balanced by bytes (encode_parallel splits by
    count, bpe.cpp:1722-1726; bytes balance better on skewed lengths)."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    if n <= 0:
        return 0, 0
    total = int(offsets[-1] - offsets[0])
    cuts = [int(np.searchsorted(offsets, offsets[0] + np.uint64(total * r // world), side="left")) for r in
            range(world + 1)]
    cuts[0], cuts[-1] = 0, n
    cuts = np.maximum.accumulate(np.minimum(cuts, n))
    return int(cuts[rank]), int(cuts[rank + 1])


def merge_word_exports(exports):
    """exports: list of (tokens uint32[T_r], offsets uint32[U_r+1], freq uint64[U_r]) per rank ->
    one concatenated (tokens, offsets, freq).  Duplicates across ranks are kept (harmless)."""
    toks, offs, freqs, base = [], [np.zeros(1, dtype=np.uint32)], [], 0
    for t, o, f in exports:
        t = np.asarray(t, dtype=np.uint32)
        o = np.asarray(o, dtype=np.uint64)
        toks.append(t)
        offs.append((o[1:] - o[0] + base).astype(np.uint32))
        freqs.append(np.asarray(f, dtype=np.uint64))
        base += int(o[-1] - o[0])
    if base >= 2 ** 32 - 16:
        raise ValueError("merged unique words exceed 2^32 tokens")
    return (np.concatenate(toks) if toks else np.zeros(0, np.uint32), np.concatenate(offs),
            np.concatenate(freqs) if freqs else np.zeros(0, np.uint64))


def all_gather_arrays(arr, group=None):
    """all_gather of variable-length 1-D numpy arrays through torch.distributed (any backend)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    raw = np.ascontiguousarray(arr).view(np.uint8)
    n = torch.tensor([raw.size], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(max(sizes), 1)
    buf = torch.zeros(m, dtype=torch.uint8, device=dev)
    if raw.size:
        buf[:raw.size] = torch.from_numpy(raw.copy()).to(dev)
    out = [torch.zeros(m, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return [o[:s].cpu().numpy().view(arr.dtype).copy() for o, s in zip(out, sizes)]


# ---------------------------------------------------------------------------------------------
# GPU paths
# ---------------------------------------------------------------------------------------------
class _DevView:
    """Zero-copy torch view of library-owned device memory (`__cuda_array_interface__`)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def encode_sharded(bpe, data, offsets, group=None, **kw):
    """Every rank encodes its contiguous sentence range on its own GPU; no collective.
    Returns (lo, hi, ids, id_offsets) for this rank."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_sentences(offsets, rank, world)
    offs = np.ascontiguousarray(np.asarray(offsets, dtype=np.uint64)[lo:hi + 1])
    ids, oo = bpe.encode_packed(data, offs, **kw) if hi > lo else (np.zeros(0, np.int32), np.zeros(1, np.uint64))
    return lo, hi, ids, oo


def train_distributed(data, model_path, vocab_size, coverage=1.0, pad_id=0, unk_id=1, bos_id=2, eos_id=3, group=None):
    """Data-parallel training over the ranks of `group` (NCCL).  `data` is the WHOLE corpus as
    bytes on every rank (each rank only uploads its own byte range).  Rank 0 writes the model."""
    import torch
    import torch.distributed as dist
    from . import _lib
    L = _lib.lib()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    pos = split_byte_ranges(data, world)
    shard = data[pos[rank]:pos[rank + 1]]
    ctx = C.c_void_p()
    if L.yttm_ctx_create(torch.cuda.current_device(), C.byref(ctx)) != 0:
        raise ValueError(L.yttm_last_error(None).decode())

    def check(rc):
        if rc != 0:
            raise ValueError(L.yttm_last_error(ctx).decode())

    try:
        check(L.yttm_train_load_corpus(ctx, C.cast(C.c_char_p(shard), C.c_void_p), len(shard), 0))
        dl, nd = C.c_uint64(0), C.c_uint64(0)
        check(L.yttm_train_char_hist(ctx, C.byref(dl), C.byref(nd)))
        # the one collective of the front-end: sum the dense code point histogram in place
        dptr, n64 = C.c_void_p(), C.c_uint64(0)
        check(L.yttm_train_char_hist_devptr(ctx, C.byref(dptr), C.byref(n64)))
        hist = torch.as_tensor(_DevView(dptr.value, n64.value, "<i8"), device="cuda")
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
        torch.cuda.synchronize()
        check(L.yttm_train_char_hist_refresh(ctx, C.byref(dl), C.byref(nd)))
        cps = np.zeros(nd.value, dtype=np.uint32)
        cnt = np.zeros(nd.value, dtype=np.uint64)
        L.yttm_train_get_char_hist(ctx, cps.ctypes.data, cnt.ctypes.data)
        char2id, n_special = choose_alphabet(cps, cnt, dl.value, coverage, (pad_id, unk_id, bos_id, eos_id))
        used = len(char2id) + n_special
        if used > vocab_size:
            raise ValueError("Incorrect arguments. Vocabulary size too small. Set vocab_size>=%d.  Current value "
                             "for vocab_size=%d" % (used, vocab_size))
        kc = np.fromiter(char2id.keys(), dtype=np.uint32)
        ki = np.fromiter(char2id.values(), dtype=np.uint32)
        check(L.yttm_train_set_alphabet(ctx, kc.ctypes.data, ki.ctypes.data, len(kc), char2id[9601]))
        st = _lib.TrainStats()
        check(L.yttm_train_build(ctx, C.byref(st)))
        # unique words of every rank -> every rank (replicated merge loop)
        nw, nt = C.c_uint64(0), C.c_uint64(0)
        check(L.yttm_train_export_words(ctx, None, 0, None, None, 0, C.byref(nw), C.byref(nt)))
        tok = np.zeros(max(nt.value, 1), dtype=np.uint32)
        off = np.zeros(nw.value + 1, dtype=np.uint32)
        frq = np.zeros(max(nw.value, 1), dtype=np.uint64)
        check(L.yttm_train_export_words(ctx, tok.ctypes.data, len(tok), off.ctypes.data, frq.ctypes.data, len(frq),
                                        C.byref(nw), C.byref(nt)))
        toks = all_gather_arrays(tok[:nt.value], group)
        offs = all_gather_arrays(off, group)
        frqs = all_gather_arrays(frq[:nw.value], group)
        mt, mo, mf = merge_word_exports(list(zip(toks, offs, frqs)))
        check(L.yttm_train_import_words(ctx, mt.ctypes.data, len(mt), mo.ctypes.data, mf.ctypes.data, len(mf),
                                        C.byref(st)))
        n_merges = vocab_size - used
        rules = np.zeros(3 * max(n_merges, 1), dtype=np.uint32)
        freqs = np.zeros(max(n_merges, 1), dtype=np.uint64)
        done = C.c_uint32(0)
        check(L.yttm_train_run(ctx, used, n_merges, rules.ctypes.data, freqs.ctypes.data, C.byref(done)))
        rules = rules[:3 * done.value].reshape(-1, 3)
        if rank == 0:
            write_model(model_path, char2id, rules, (pad_id, unk_id, bos_id, eos_id), vocab_size)
        dist.barrier(group)
        return int(done.value)
    finally:
        L.yttm_ctx_destroy(ctx)


def choose_alphabet(cps, counts, data_len, coverage, special):
    """compute_alphabet_helper (bpe.cpp:316-355) in numpy/python; returns ({cp: internal id}, n_special)."""
    order = sorted(zip(counts.tolist(), cps.tolist()))
    cur = removed = 0
    while cur < len(order) and float(data_len - removed - order[cur][0]) > float(data_len) * coverage:
        removed += order[cur][0]
        cur += 1
    n_special = sum(1 for s in special if s != -1)
    char2id = {9601: n_special}
    nxt = n_special + 1
    for c, cp in reversed(order[cur:]):
        char2id[cp] = nxt
        nxt += 1
    return char2id, n_special


def write_model(path, char2id, rules, special, vocab_size):
    """rename_tokens (bpe.cpp:814-837) + BPEState::dump (utils.cpp:50-66); the char2id lines come in the
    reference's flat_hash_map iteration order (yttm_api_dump_order), so the file is byte-identical."""
    from . import _lib
    pad, unk, bos, eos = special
    taken = {s for s in special if s != -1}
    n_special = len(taken)
    ren = {}
    cur = n_special
    for i in range(vocab_size):
        if i not in taken:
            ren[cur] = i
            cur += 1
    filled = np.array(sorted(char2id, key=lambda cp: char2id[cp]), dtype=np.uint32)
    order = np.zeros(len(filled), dtype=np.uint32)
    L = _lib.lib()
    if L.yttm_api_dump_order(filled.ctypes.data, len(filled), order.ctypes.data) != 0:
        raise ValueError("duplicate code points in char2id")
    with open(path, "w") as f:
        f.write("%d %d\n" % (len(char2id), len(rules)))
        for cp in order.tolist():
            f.write("%d %d\n" % (cp, ren[char2id[cp]]))
        for x, y, z in rules:
            f.write("%d %d %d\n" % (ren[int(x)], ren[int(y)], ren[int(z)]))
        f.write("%d %d %d %d\n" % (unk, pad, bos, eos))
