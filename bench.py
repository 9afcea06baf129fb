#!/usr/bin/env python
"""bench.py — headline measurement of the BPE hot paths on B200 (contract: task statement ④).

One "step" = one pass of hot path (b), batch encode_as_ids, over BASELINE.json configs[1]:
1 M synthetic 128-byte sentences, vocab 32 000 (model trained by this framework's own GPU trainer
on synthetic text of the same distribution, outside the timed region).

  value   Msent/s, whole job, inputs already resident in HBM (yttm_enc_run_device)
  e2e     same metric through the host-buffer C-ABI call (yttm_enc_run): pinned host input,
          H2D + kernels + D2H of ids/offsets inside the timed region
  roofline             dominant kernel of the step (encode_words_kernel), algorithmic bytes
                       sum(len_i + 4 n_ids_i + 16) / its CUDA-event duration
  roofline_train_scan  the pair-count scan of hot path (a) on a packed token buffer >> L2
                       (4T + 12U bytes per launch), the kernel BASELINE.json's 70 % target names
  train                sizes / stage times of the GPU training run that produced the model
  cpu_baseline         the unmodified reference (oracle/_ref prod build) encode_as_ids on the host

`--impl reference` times the reference's own CPU implementation (all host threads) on the same
workload.  N > 1 (torchrun): sentences shard by rank, no collective on the data path (weak scaling).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CACHE = os.environ.get("YTTM_BENCH_CACHE", "/tmp/yttm_b200_bench_cache")
N_SENT, SENT_LEN, VOCAB = 1_000_000, 128, 32_000
TRAIN_BYTES = 100_000_000


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError, TypeError):  # absent or of another shape: the recipe's stated fallback
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def workload(rank, n_sent, train_bytes):
    """(train text, sentence bytes, uint64 offsets) — cached on local disk between the two arms."""
    from youtokentome_b200 import synth
    os.makedirs(CACHE, exist_ok=True)
    tp = os.path.join(CACHE, "train_%d.bin" % train_bytes)
    sp = os.path.join(CACHE, "sent_%d_%d_r%d" % (n_sent, SENT_LEN, rank))
    fz = None
    if os.path.exists(tp):
        text = open(tp, "rb").read()
    else:
        fz = synth.FastZipf(n_words=200_000, s=1.07, seed=1234)
        text = fz.text(train_bytes)
        if rank == 0:
            with open(tp + ".tmp%d" % os.getpid(), "wb") as f:
                f.write(text)
            os.replace(tp + ".tmp%d" % os.getpid(), tp)
    if os.path.exists(sp + ".bin") and os.path.exists(sp + ".npy"):
        buf = open(sp + ".bin", "rb").read()
        offs = np.load(sp + ".npy")
    else:
        fz = fz or synth.FastZipf(n_words=200_000, s=1.07, seed=1234)
        buf, offs = fz.packed_sentences(n_sent, SENT_LEN, seed=4321 + rank)
        with open(sp + ".bin.tmp", "wb") as f:
            f.write(buf)
        os.replace(sp + ".bin.tmp", sp + ".bin")
        np.save(sp + ".npy", offs)
    return text, buf, offs


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.rows, self.proc = [], None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_encode(model, buf, offs, cores, max_sent):
    """The unmodified reference (prod build) encode_as_ids with `cores` threads on a bounded sample."""
    import _bind
    kind = "reference" if _bind.have_reference("prod") else "port"
    n = min(len(offs) - 1, max_sent)
    o = np.ascontiguousarray(offs[:n + 1])
    if kind == "reference":
        enc = _bind.Reference("prod").encoder(model, n_threads=cores)
        ids, _ = enc.encode_packed(buf, o)
        sec = enc.last_seconds
    else:
        enc = _bind.Oracle().encoder(model)
        ids, _ = enc.encode_packed(buf, o)
        sec, cores = enc.last_seconds, 1
    return {"value": n / sec / 1e6, "unit": "Msent/s", "cores": cores, "kind": kind,
            "sample": "%d of the %d sentences (%.0f MB), encode_as_ids only, %.2f s" %
                      (n, len(offs) - 1, float(o[-1]) / 1e6, sec)}, ids


def reference_arm(args, rank, world):
    """--impl reference: the reference's CPU encode_as_ids on the same workload, all host threads."""
    if rank != 0:
        return
    import _bind
    _bind.build_checkers()
    cores = os.cpu_count() or 1
    text, buf, offs = workload(0, N_SENT, TRAIN_BYTES)
    model = os.path.join(CACHE, "model_ref_%d.yttm" % VOCAB)
    if not os.path.exists(model):
        # the model must equal the GPU arm's: train it with the reference itself (DETERMINISTIC_QUEUE
        # build = the tie-break order both implementations are pinned to)
        kind = "det" if _bind.have_reference("det") else None
        if kind:
            _bind.Reference("det").train(text, model, VOCAB, 1.0, n_threads=min(8, cores))
        else:
            _bind.Oracle().train(text, model, VOCAB, 1.0)
    times = []
    sample = min(N_SENT, max(50_000, 40_000 * cores))
    base = None
    for it in range(args.warmup + args.steps):
        base, _ = cpu_reference_encode(model, buf, offs, cores, sample)
        if it >= args.warmup:
            times.append(sample / (base["value"] * 1e6))
    sec = sum(times) / len(times)
    val = sample / sec / 1e6
    base["value"] = val
    out = {"impl": "reference", "metric": "encode throughput, 1M x 128 B synthetic sentences, vocab 32k",
           "value": val, "unit": "Msent/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "u8/u32", "data": "synthetic",
           "config": {"workload": "configs[1]: encode 1M synthetic 128-byte sentences, vocab 32k",
                      "step": "reference encode_as_ids on %d sentences, %d threads" % (sample, cores)},
           "cpu_baseline": base,
           "e2e": {"value": val, "unit": "Msent/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


_REAL_STDOUT = None


def experimental_ab(model, budget_s=300.0):
    """INFORMATIONAL, rank 0 at N = 1 only, after every measurement of the line has been taken: the env-gated
    experimental kernels (off by default, parity-checked under the CPU emulator) against the default ones
    on this box, each in its own subprocess with a hard timeout so that neither a wrong nor a hanging experimental
    kernel can cost the JSON line.  Nothing here enters `value`, `e2e` or `roofline`."""
    import tempfile
    deadline = time.time() + budget_s   # the whole leg, however many of its subprocesses hang
    left = lambda cap: max(1.0, min(cap, deadline - time.time()))
    env = {k: v for k, v in os.environ.items()
           if not k.startswith(("YTTM_ENC_", "YTTM_LOOP_", "YTTM_DBG")) and k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    out = {"note": "informational A/B of env-gated experimental kernels vs the defaults (median CUDA-event ms per stage, "
                   "same workload and model as the line); not part of value / e2e"}
    try:
        with tempfile.TemporaryDirectory() as d:
            js = os.path.join(d, "ab.json")
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ab_encode.py"), str(N_SENT), "5", js, model],
                               env=env, cwd=ROOT, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=left(150))
            if r.returncode == 0 and os.path.exists(js):
                with open(js) as fh:
                    out["encode_stage_ms"] = json.load(fh)
            else:
                out["encode_stage_ms"] = {"error": r.stderr.decode(errors="replace")[-300:]}
    except subprocess.TimeoutExpired:
        out["encode_stage_ms"] = {"error": "timeout"}
    except Exception as e:
        out["encode_stage_ms"] = {"error": repr(e)}
    loop = {}
    for name, extra in (("default", {}), ("wide_probe", {"YTTM_LOOP_WIDEPROBE": "1"}),
                        ("max_load_50", {"YTTM_PAIR_MAX_LOAD_PCT": "50"}),
                        ("wide_probe+max_load_50", {"YTTM_LOOP_WIDEPROBE": "1", "YTTM_PAIR_MAX_LOAD_PCT": "50"}),
                        ("per_block_timers", {"YTTM_DBG": "8"}), ("wide_probe+per_block_timers", {"YTTM_LOOP_WIDEPROBE": "1", "YTTM_DBG": "8"}),
                        ("pinned_h2d_8_threads", {"YTTM_TRAIN_PINNED_H2D": "8"}),
                        ("blocks_74", {"YTTM_LOOP_BLOCKS": "74"}), ("blocks_111", {"YTTM_LOOP_BLOCKS": "111"}),
                        ("threads_512", {"YTTM_LOOP_THREADS": "512"}), ("threads_256", {"YTTM_LOOP_THREADS": "256"})):
        if deadline - time.time() < 15:
            loop[name] = {"error": "skipped: the leg's %d s budget is spent" % budget_s}
            continue
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "probe_train.py"), "zipf", str(VOCAB), str(TRAIN_BYTES)],
                               env=dict(env, **extra), cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=left(60))
            last = [ln for ln in r.stdout.decode(errors="replace").splitlines() if ln.startswith("{")]
            if r.returncode == 0 and last:
                j = json.loads(last[-1])
                loop[name] = {"us_per_merge": j["us_per_merge"], "merges": j["merges"], "table_slots": j["cap"], "launches": j["launches"],
                              "h2d_ms": j["front_ms"]["h2d"], "phase_us_per_iter": j["phase_us_per_iter"]}
            else:
                loop[name] = {"error": "rc %d" % r.returncode}
        except subprocess.TimeoutExpired:
            loop[name] = {"error": "timeout"}
        except Exception as e:
            loop[name] = {"error": repr(e)}
    out["merge_loop"] = loop
    return out


def quiet_stdout():
    """stdout must carry the ONE JSON line and nothing else, but libraries write to fd 1 behind
    Python's back (NCCL prints its version banner there): point fd 1 at stderr for the whole run
    and keep a private duplicate of the real stdout for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(out):
    line = (json.dumps(out) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, line)
    else:
        os.write(_REAL_STDOUT, line)


T_START = time.time()


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-experimental-ab", action="store_true")
    ap.add_argument("--scan-tokens", type=int, default=256 * 1024 * 1024)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from youtokentome_b200 import _lib
    from _gpu import gpu_train
    torch.cuda.set_device(local)
    os.environ["YTTM_DEVICE"] = str(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    L = _lib.lib()
    hbm_peak, peak_src = peaks()

    text, buf, offs = workload(rank, N_SENT, TRAIN_BYTES)
    n_sent, n_bytes = len(offs) - 1, int(offs[-1])

    # ---- model: this framework's GPU trainer (hot path a), outside the timed region
    model = os.path.join(CACHE, "model_gpu_%d_r%d.yttm" % (VOCAB, rank))
    def read_report():
        rep = (C.c_double * 16)()
        L.yttm_api_train_report(rep, 16)
        names = ["n_bytes", "data_len", "n_words", "n_unique", "n_tokens", "n_pairs", "n_merges", "read_s", "h2d_ms",
                 "char_hist_ms", "word_count_ms", "tokenise_ms", "pair_hist_ms", "merge_loop_ms", "total_s", "launches"]
        return dict(zip(names, [float(x) for x in rep]))

    t0 = time.perf_counter()
    gpu_train(text, VOCAB, 1.0, model=model)           # cold: first CUDA work of the process, clocks still ramping
    cold_wall = time.perf_counter() - t0
    runs = []
    for _ in range(3):                                 # warm runs; the median is reported (the box is shared: noisy)
        t0 = time.perf_counter()
        gpu_train(text, VOCAB, 1.0, model=model)
        runs.append((time.perf_counter() - t0, read_report()))
    runs.sort(key=lambda r: r[0])
    train_wall, train = runs[1]
    train["wall_s"] = train_wall
    train["cold_wall_s"] = cold_wall
    train["warm_walls_s"] = [r[0] for r in runs]
    train["merge_loop_ms_runs"] = [r[1]["merge_loop_ms"] for r in runs]
    train["GBps_e2e"] = len(text) / train_wall / 1e9
    train["us_per_merge"] = train["merge_loop_ms"] * 1e3 / max(train["n_merges"], 1)

    # ---- encoder handle + device-resident inputs
    h = L.yttm_api_open(model.encode(), 1)
    assert h, L.yttm_api_last_error(None)
    ctx, enc = L.yttm_api_device_context(h), L.yttm_api_device_encoder(h)
    host_bytes = torch.frombuffer(bytearray(buf), dtype=torch.uint8).pin_memory()
    host_offs = torch.from_numpy(offs.astype(np.int64)).pin_memory()
    d_bytes, d_offs = host_bytes.cuda(), host_offs.cuda()
    out_cap = n_bytes + 3 * n_sent + 16
    host_ids = torch.empty(out_cap, dtype=torch.int32).pin_memory()
    host_oo = torch.empty(n_sent + 1, dtype=torch.int64).pin_memory()

    def step_device():
        out_n = C.c_uint64(0)
        p1, p2 = C.c_void_p(), C.c_void_p()
        rc = L.yttm_enc_run_device(enc, d_bytes.data_ptr(), d_offs.data_ptr(), n_bytes, n_sent, 0, 0, 0, 0.0, 0, 0,
                                   C.byref(p1), C.byref(p2), C.byref(out_n))
        assert rc == 0, L.yttm_last_error(ctx)
        return out_n.value

    def step_host():
        out_n = C.c_uint64(0)
        rc = L.yttm_enc_run(enc, host_bytes.data_ptr(), host_offs.data_ptr(), n_sent, 0, 0, 0, 0.0, 0, 0,
                            host_ids.data_ptr(), out_cap, host_oo.data_ptr(), C.byref(out_n))
        assert rc == 0, L.yttm_last_error(ctx)
        return out_n.value

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            n_ids = fn()
        barrier()
        l0 = L.yttm_launch_count(ctx)
        kern = {"enc_find": 0.0, "enc_words": 0.0, "enc_gather": 0.0, "enc_scan": 0.0}
        t0 = time.perf_counter()
        for _ in range(steps):
            n_ids = fn()
            for k in kern:
                kern[k] += L.yttm_stage_ms(ctx, k.encode())
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        launches = L.yttm_launch_count(ctx) - l0
        if world > 1:
            t = torch.tensor([sec], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        return sec, n_ids, {k: v / steps for k, v in kern.items()}, launches

    # clocks / throttle reasons are sampled (100 ms period) from before the warm-up of the first timed
    # loop to the end of the second one: the timed regions themselves last only tens of milliseconds
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.3 if rank == 0 else 0.0)
    sec_d, n_ids, kern, launches = timed(step_device, args.steps, args.warmup)
    sec_h, n_ids_h, _, _ = timed(step_host, args.steps, args.warmup)
    for _ in range(20):                       # keep the GPU busy long enough for a few more samples
        step_device()
    clocks = sampler.stop() if sampler else None
    assert n_ids == n_ids_h

    value = world * n_sent * args.steps / sec_d / 1e6
    e2e = world * n_sent * args.steps / sec_h / 1e6
    algo = n_bytes + 4 * n_ids + 16 * n_sent
    dom = max(kern, key=lambda k: kern[k])
    ach = algo / (kern[dom] * 1e-3) / 1e9 if kern[dom] > 0 else None
    roofline = {"bound": "hbm", "kernel": {"enc_words": "encode_words_kernel", "enc_find": "find_words_kernel",
                                           "enc_gather": "gather_ids_kernel", "enc_scan": "scan"}[dom],
                "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak if ach else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of ONE ncu --set full capture of this kernel on
                # this workload (profiles/r01_prof_encode_words.raw.csv: 591.2 MB + 468.0 MB); not re-measured here
                "traffic": 1059194624 if dom == "enc_words" and n_sent == N_SENT else None,
                "traffic_source": "profiles/r01_prof_encode_words.raw.csv (ncu --set full, same workload)",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": algo, "kernel_ms": kern}

    # ---- hot path (a): the per-merge-iteration scan of the packed token buffer, on a buffer >> L2
    # (rank 0 only).  Each iteration of merge_loop_kernel streams every token slot and word offset
    # once (TMA-staged tiles): algorithmic bytes 4T + 4U per iteration (frequencies are only read
    # for rewritten words).  Time = device-side timers of the apply phase incl. its tail barrier.
    scan = None
    if rank == 0 and args.scan_tokens > 0:
        def scan_probe(log2_first):
            """12 merges over synthetic words of 8 tokens; log2_first = 0: all words start with the same
            token, every merge rewrites ~16 700 words (mid-training regime); 10: 1024 different initial
            tokens, a merge rewrites a few dozen words (the bulk of a long training run)."""
            c2 = C.c_void_p()
            assert L.yttm_ctx_create(local, C.byref(c2)) == 0
            wl, alpha, iters = 8, 2000, 12
            n_w = args.scan_tokens // wl
            rc = L.yttm_train_synth_words(c2, n_w, wl, alpha | (log2_first << 24), 7)
            assert rc == 0, L.yttm_last_error(c2)
            rules = np.zeros(3 * iters, dtype=np.uint32)
            fr = np.zeros(iters, dtype=np.uint64)
            nd = C.c_uint32(0)
            assert L.yttm_train_run(c2, 4 + (1 << log2_first) + alpha, iters, rules.ctypes.data, fr.ctypes.data,
                                    C.byref(nd)) == 0, L.yttm_last_error(c2)
            g = lambda k: L.yttm_stage_ms(c2, k.encode())
            it = max(g("loop_iters"), 1.0)
            t_scan = (g("loop_apply") + g("loop_barrier2")) / it          # ms per merge
            ab = 4 * args.scan_tokens + 4 * n_w
            ach = ab / (t_scan * 1e-3) / 1e9
            out = {"achieved": ach, "frac": ach / hbm_peak, "ms_per_iteration_scan": t_scan,
                   "iterations": int(nd.value), "resident": int(g("loop_resident")),
                   "words_rewritten_per_merge": float(fr[:nd.value].mean() / 4.5) if nd.value else None,
                   "phase_ms_per_iter": {k: g(k) / it for k in ["loop_argmax", "loop_barrier1", "loop_apply",
                                                                 "loop_barrier2"]},
                   "table_slots": g("table_capacity")}
            ms, ab2 = C.c_double(0), C.c_uint64(0)
            L.yttm_train_scan_once(c2, C.byref(ms), C.byref(ab2))
            out["initial_histogram"] = {"kernel": "pair_hist_kernel", "ms": ms.value,
                                        "GBps": ab2.value / (ms.value * 1e-3) / 1e9}
            L.yttm_ctx_destroy(c2)
            return out, ab, n_w
        heavy, ab, n_w = scan_probe(0)
        light, _, _ = scan_probe(10)
        scan = {"bound": "hbm", "kernel": "merge_loop_kernel (apply phase, STREAMING tiles through the TMA ring)",
                "achieved": heavy["achieved"], "peak": hbm_peak, "unit": "GB/s", "frac": heavy["frac"],
                "traffic": None,
                # the ncu capture is of a smaller probe (one launch = 6 merges over 64 Mi tokens, arg-max sweeps
                # of the 16 Mi-slot table included), so it is reported beside, not as, this launch's traffic
                "traffic_ncu": {"capture": "profiles/r01_prof_merge_loop_stream_v4.raw.csv", "tokens": 67108864,
                                "merges": 6, "dram_bytes": 2914580920, "algorithmic_bytes_scan": 6 * 301989888,
                                "algorithmic_bytes_argmax_sweep": 6 * 134217728},
                "algorithmic_bytes_per_launch": ab, "tokens": args.scan_tokens, "words": n_w,
                "peak_source": peak_src, "heavy_merges": heavy, "light_merges": light}

    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        import _bind
        cores = os.cpu_count() or 1
        cpu, ref_ids = cpu_reference_encode(model, buf, offs, cores, min(n_sent, max(50_000, 40_000 * cores)))
        # parity spot check of the timed configuration (not timed): GPU ids == reference ids
        step_host()
        k = len(ref_ids)
        assert np.array_equal(host_ids[:k].numpy(), ref_ids), "bench: GPU ids differ from the reference"
        cpu["ids_equal_on_sample"] = True
        # the reference's trainer on the same corpus, 8 threads (its hard cap, bpe.cpp:1348), for `train`
        if _bind.have_reference("prod"):
            sec = _bind.Reference("prod").train(text, os.path.join(CACHE, "model_refprod.yttm"), VOCAB, 1.0,
                                                n_threads=min(8, cores))
            train["cpu_reference"] = {"seconds": sec, "GBps": len(text) / sec / 1e9, "threads": min(8, cores),
                                      "kind": "reference", "sample": "learn_bpe_from_string on the same %d MB"
                                                                     % (len(text) // 1_000_000)}
            train["speedup_vs_cpu_reference_wall"] = sec / train["wall_s"]
        # SURVEY.md §8d also asks for the single-thread figures (bounded: 50 k sentences / one training run)
        try:
            one, _ = cpu_reference_encode(model, buf, offs, 1, min(n_sent, 50_000))
            cpu["one_thread"] = {"value": one["value"], "unit": "Msent/s", "sample": one["sample"]}
            if _bind.have_reference("prod"):
                sec1 = _bind.Reference("prod").train(text, os.path.join(CACHE, "model_refprod1.yttm"), VOCAB, 1.0,
                                                     n_threads=1)
                train["cpu_reference_1thread"] = {"seconds": sec1, "GBps": len(text) / sec1 / 1e9, "threads": 1,
                                                  "kind": "reference"}
        except Exception as e:  # the extra legs must never cost the JSON line
            cpu["one_thread"] = {"error": repr(e)}

    ab = None
    if rank == 0 and args.gpus == 1 and world == 1 and not args.no_cpu_baseline and not args.no_experimental_ab:
        if time.time() - T_START < 420:   # a slow box: the line matters more than the extras
            ab = experimental_ab(model)
        else:
            ab = {"note": "skipped: the run had already taken %.0f s" % (time.time() - T_START)}

    if ab is not None and "encode_stage_ms" in ab:
        # e2e with a smaller FIRST chunk (host-only knob of yttm_enc_run, default kernels; DESIGN.md knob table): the
        # copy-in of the first chunk overlaps nothing.  Same timed() as the line's e2e, after it, informational.
        try:
            e2e_ab = {}
            for mb in ("8", "16", None):
                if mb:
                    os.environ["YTTM_ENC_FIRST_CHUNK_MB"] = mb
                else:
                    os.environ.pop("YTTM_ENC_FIRST_CHUNK_MB", None)
                sec_x, _, _, _ = timed(step_host, args.steps, 2)
                e2e_ab["first_chunk_%s_mb" % mb if mb else "default_again"] = n_sent * args.steps / sec_x / 1e6
            ab["e2e_msent_per_s"] = e2e_ab
        except Exception as e:
            ab["e2e_msent_per_s"] = {"error": repr(e)}
        finally:
            os.environ.pop("YTTM_ENC_FIRST_CHUNK_MB", None)

    if rank == 0:
        out = {"metric": "encode throughput, 1M x 128 B synthetic sentences, vocab 32k", "value": value,
               "unit": "Msent/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": sec_d / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "u8/u32", "data": "synthetic",
               "config": {"workload": "configs[1]: encode 1M synthetic 128-byte sentences, vocab 32k, per GPU",
                          "sentences_per_gpu": n_sent, "bytes_per_gpu": n_bytes, "ids_per_gpu": n_ids,
                          "l2": "inputs + slot buffers (%.0f MB) exceed the 126 MB L2" % ((5 * n_bytes) / 1e6),
                          "parallelism": "sentences sharded by rank, no collective"},
               "e2e": {"value": e2e, "unit": "Msent/s", "h2d_bytes_per_step": n_bytes + 8 * (n_sent + 1),
                       "d2h_bytes_per_step": 4 * n_ids + 8 * (n_sent + 1), "ms_per_step": sec_h / args.steps * 1e3},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
               "roofline_train_scan": scan, "train": train, "cpu_baseline": cpu}
        if ab is not None:
            out["experimental_ab"] = ab
        emit(out)
    L.yttm_api_close(h)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
