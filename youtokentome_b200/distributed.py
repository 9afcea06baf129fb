"""One-process-per-GPU plumbing (torch.distributed) around the C ABI — SURVEY.md §8e.

Encode shards by sentence: contiguous ranges balanced by bytes, NO collective on the data path
(`shard_sentences`, `encode_sharded`).

Training shards the corpus by byte range cut at ASCII spaces exactly as the reference cuts it
between threads (bpe.cpp:864-873); see train_distributed.
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


def check_config(vocab_size, coverage, pad_id, unk_id, bos_id, eos_id):
    """check_config of the reference (bpe.cpp:1295-1350), same texts; raises ValueError."""
    if coverage <= 0 or coverage > 1:
        raise ValueError("coverage value must be in the range (0, 1]. Current value of coverage = %f" % coverage)
    if unk_id < 0 or unk_id >= vocab_size:
        raise ValueError("unk_id: must be in the range [0, vocab_size - 1]. Current value of vocab_size = %d; unk_id = %d"
                         % (vocab_size, unk_id))
    for name, v, sep in (("pad_id", pad_id, ";"), ("bos_id", bos_id, ";"), ("eos_id", eos_id, "")):
        if v < -1 or v >= vocab_size:
            raise ValueError("%s must be in the range [-1, vocab_size - 1]. Current value of vocab_size = %d%s %s = %d"
                             % (name, vocab_size, sep, name, v))
    ids = [unk_id] + [v for v in (pad_id, bos_id, eos_id) if v != -1]
    if len(set(ids)) != len(ids):
        raise ValueError("All ids of special tokens must be different.")


class LocalComm:
    """world = 1: the same code path as a multi-GPU job without torch.distributed (bench.py at --gpus 1, tests)."""
    rank, world = 0, 1

    def all_gather_bytes(self, b):
        return bytes(b)

    def allreduce_sum_u64(self, ptr, n):
        pass

    def agree(self, ok):
        return bool(ok)

    def all_to_all(self, ptr, counts, itemsize):
        return None, ptr, [int(counts[0])]

    def barrier(self):
        pass


class TorchComm:
    """The collectives train_distributed needs, over torch.distributed (NCCL on the GPUs; plumbing only — the per-merge
    exchange of the merge loop is NOT here, it is peer stores inside the kernel, csrc/merge_loop.cuh)."""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")

    def all_gather_bytes(self, b):
        t = self.torch.frombuffer(bytearray(b), dtype=self.torch.uint8).to(self.dev)
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t, group=self.group)
        return b"".join(bytes(o.cpu().numpy().tobytes()) for o in out)

    def allreduce_sum_u64(self, ptr, n):
        t = self.torch.as_tensor(_DevView(ptr, n, "<i8"), device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        self.torch.cuda.synchronize()

    def agree(self, ok):
        """False on every rank if any rank says False (a failure must not leave the others inside a collective)."""
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
        return bool(t.item())

    def all_to_all(self, ptr, counts, itemsize):
        """counts[d] items of `itemsize` bytes go to rank d (device buffer at ptr, destinations back to back).
        Returns (received tensor kept alive by the caller, its device pointer, counts per source)."""
        torch, dist = self.torch, self.dist
        send_n = torch.tensor([int(c) for c in counts], dtype=torch.int64, device=self.dev)
        recv_n = torch.zeros_like(send_n)
        dist.all_to_all_single(recv_n, send_n, group=self.group)
        recv = [int(x) for x in recv_n.tolist()]
        total = int(sum(int(c) for c in counts)) * itemsize
        src = torch.as_tensor(_DevView(ptr, max(total, 1), "|u1"), device=self.dev)[:total]
        out = torch.empty(max(sum(recv) * itemsize, 1), dtype=torch.uint8, device=self.dev)
        dist.all_to_all_single(out[:sum(recv) * itemsize], src, [r * itemsize for r in recv],
                               [int(c) * itemsize for c in counts], group=self.group)
        torch.cuda.synchronize()
        return out, out.data_ptr(), recv

    def barrier(self):
        self.dist.barrier(self.group)


def train_distributed(data, model_path, vocab_size, coverage=1.0, pad_id=0, unk_id=1, bos_id=2, eos_id=3, group=None,
                      comm=None, device=None, lib=None, sharded=False, stats_out=None):
    """Training over the ranks of `group`, one process per GPU.  `data` is the WHOLE corpus as bytes on every rank (each
    rank uploads only its own byte range, cut as the reference cuts it between threads, bpe.cpp:864-873) or, with
    sharded=True, this rank's shard (which must end at a word boundary).  Every rank returns the number of merges;
    rank 0 writes the model.

    Front end on every GPU: code point histogram (one NCCL allreduce), word split + dedup of the shard, then the unique
    words are hash-partitioned across the ranks (all-to-all of device buffers) so that equal words of different shards
    meet on one rank and are counted once — the job-wide set of unique words is cut into `world` disjoint parts, the
    reference's thread partition (bpe.cpp:1066-1069) and its merge of the per-thread word maps (:1029-1039).
    Merge loop: every rank rewrites its own words; the count changes of a merge go straight into every peer's exchange
    buffer (peer stores over NVLink inside the persistent kernel, csrc/merge_loop.cuh); every rank keeps the full pair
    table and elects the same pair.  No collective call per merge."""
    from . import _lib
    L = lib or _lib.lib()
    check_config(vocab_size, coverage, pad_id, unk_id, bos_id, eos_id)
    comm = comm or TorchComm(group)
    rank, world = comm.rank, comm.world
    if world > 8:
        raise ValueError("train_distributed: at most 8 ranks (one NVLink domain)")
    if device is None:
        import torch
        device = torch.cuda.current_device()
    if sharded:
        shard = data
    else:
        pos = split_byte_ranges(data, world)
        shard = data[pos[rank]:pos[rank + 1]]
    ctx = C.c_void_p()
    ok = L.yttm_ctx_create(device, C.byref(ctx)) == 0
    err = None if ok else L.yttm_last_error(None).decode()
    if not comm.agree(ok):
        raise ValueError(err or "train_distributed: another rank failed to create its CUDA context")
    keep = []  # received tensors stay alive until the import has copied them
    import time
    walls, t_last = {}, [time.perf_counter()]

    def lap(name):   # host wall clock between the steps of the protocol (stats_out["host_ms"])
        now = time.perf_counter()
        walls[name] = walls.get(name, 0.0) + (now - t_last[0]) * 1e3
        t_last[0] = now

    def step(fn):
        """run one local phase; every rank learns whether all of them succeeded before the next collective"""
        msg = None
        try:
            rc = fn()
            if rc:
                msg = L.yttm_last_error(ctx).decode()
        except ValueError as e:
            msg = str(e)
        if not comm.agree(msg is None):
            raise ValueError(msg or "train_distributed: another rank failed")

    try:
        handle = C.create_string_buffer(128)
        step(lambda: L.yttm_train_dist_init(ctx, rank, world, handle))
        handles = comm.all_gather_bytes(handle.raw)
        step(lambda: L.yttm_train_dist_connect(ctx, handles))
        lap("ctx+exchange_buffer")
        dl, nd = C.c_uint64(0), C.c_uint64(0)
        step(lambda: L.yttm_train_load_corpus(ctx, C.cast(C.c_char_p(shard), C.c_void_p), len(shard), 0) or
             L.yttm_train_char_hist(ctx, C.byref(dl), C.byref(nd)))
        lap("load_corpus+char_hist")
        # the one collective of the byte passes: sum the dense code point histogram in place
        dptr, n64 = C.c_void_p(), C.c_uint64(0)
        step(lambda: L.yttm_train_char_hist_devptr(ctx, C.byref(dptr), C.byref(n64)))
        comm.allreduce_sum_u64(dptr.value, n64.value)
        step(lambda: L.yttm_train_char_hist_refresh(ctx, C.byref(dl), C.byref(nd)))
        cps = np.zeros(nd.value, dtype=np.uint32)
        cnt = np.zeros(nd.value, dtype=np.uint64)
        L.yttm_train_get_char_hist(ctx, cps.ctypes.data, cnt.ctypes.data)
        char2id, n_special = choose_alphabet(cps, cnt, dl.value, coverage, (pad_id, unk_id, bos_id, eos_id))
        used = len(char2id) + n_special
        if used > vocab_size:  # identical on every rank (same histogram): no agreement round needed
            raise ValueError("Incorrect arguments. Vocabulary size too small. Set vocab_size>=%d.  Current value "
                             "for vocab_size=%d" % (used, vocab_size))
        kc = np.fromiter(char2id.keys(), dtype=np.uint32)
        ki = np.fromiter(char2id.values(), dtype=np.uint32)
        lap("allreduce+alphabet")
        nu = C.c_uint64(0)
        step(lambda: L.yttm_train_set_alphabet(ctx, kc.ctypes.data, ki.ctypes.data, len(kc), char2id[9601]) or
             L.yttm_train_dist_word_table(ctx, C.byref(nu)))
        lap("word_table")
        # unique words -> their owner ranks (device buffers, all-to-all)
        bpd, wpd = (C.c_uint64 * 8)(), (C.c_uint64 * 8)()
        p_b, p_p, p_f = C.c_void_p(), C.c_void_p(), C.c_void_p()
        step(lambda: L.yttm_train_dist_export_words(ctx, bpd, wpd, C.byref(p_b), C.byref(p_p), C.byref(p_f)))
        t_b, r_b, bps = comm.all_to_all(p_b.value, list(bpd)[:world], 1)
        t_p, r_p, wps = comm.all_to_all(p_p.value, list(wpd)[:world], 8)
        t_f, r_f, _ = comm.all_to_all(p_f.value, list(wpd)[:world], 8)
        keep += [t_b, t_p, t_f]
        lap("word_exchange")
        st = _lib.TrainStats()
        bsrc, wsrc = (C.c_uint64 * 8)(*bps), (C.c_uint64 * 8)(*wps)
        step(lambda: L.yttm_train_dist_import_words(ctx, r_b, bsrc, r_p, r_f, wsrc, C.byref(st)))
        lap("import+tokenise+pair_table")
        n_merges = vocab_size - used
        rules = np.zeros(3 * max(n_merges, 1), dtype=np.uint32)
        freqs = np.zeros(max(n_merges, 1), dtype=np.uint64)
        done = C.c_uint32(0)
        step(lambda: L.yttm_train_run(ctx, used, n_merges, rules.ctypes.data, freqs.ctypes.data, C.byref(done)))
        rules = rules[:3 * done.value].reshape(-1, 3)
        lap("merge_loop")
        stats = {"n_unique": int(st.n_unique), "n_tokens": int(st.n_tokens), "n_pairs": int(st.n_pairs),
                                  "n_merges": int(done.value), "merge_loop_ms": L.yttm_stage_ms(ctx, b"merge_loop"),
                                  "launches": L.yttm_stage_ms(ctx, b"loop_launches"),
                                  "front_ms": {k: L.yttm_stage_ms(ctx, k.encode()) for k in
                                               ("h2d", "char_hist", "word_count", "word_import", "tokenise", "pair_hist")},
                                  "phase_us_per_iter": {k: L.yttm_stage_ms(ctx, k.encode()) * 1e3 /
                                                        max(L.yttm_stage_ms(ctx, b"loop_iters"), 1.0) for k in
                                                        ("loop_elect", "loop_apply", "loop_partition", "loop_drain")}}
        train_distributed.last = stats  # one process per rank; threads of a test harness pass stats_out instead
        if stats_out is not None:
            stats_out.update(stats)
        if rank == 0 and model_path:
            write_model(model_path, char2id, rules, (pad_id, unk_id, bos_id, eos_id), vocab_size, lib=L)
        comm.barrier()
        lap("write_model")
        if stats_out is not None:
            stats_out["host_ms"] = {k: round(v, 2) for k, v in walls.items()}
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


def write_model(path, char2id, rules, special, vocab_size, lib=None):
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
    L = lib or _lib.lib()
    if L.yttm_api_dump_order(filled.ctypes.data, len(filled), order.ctypes.data) != 0:
        raise ValueError("duplicate code points in char2id")
    with open(path, "w") as f:
        f.write("%d %d\n" % (len(char2id), len(rules)))
        for cp in order.tolist():
            f.write("%d %d\n" % (cp, ren[char2id[cp]]))
        for x, y, z in rules:
            f.write("%d %d %d\n" % (ren[int(x)], ren[int(y)], ren[int(z)]))
        f.write("%d %d %d %d\n" % (unk, pad, bos, eos))
