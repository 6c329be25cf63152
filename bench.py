#!/usr/bin/env python3
"""bench.py -- throughput (GB/s of input scanned) of the matching hot path on synthetic inputs, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4|c5] [--impl reference]

A "step" is one pass of the device path over one resident batch.  Workloads (BASELINE.json configs):
  c3 (default)  configs[2]: URLCapture TDFA, FindAllBytes over 4 GiB of synthetic prose per GPU -- the configuration the
                north_star's target is quoted on.  N > 1: ONE logical buffer of N x 4 GiB, sharded, cursors exchanged.
  c2            configs[1]: email backtracking pattern, FindAllBytes over a 1 GiB synthetic log buffer.
  c4            configs[3]: the corpus pattern suite, batched MatchBytes over 12.5 M short inputs per GPU (100 M on 8),
                one multi-program launch; input-sharded, flags gathered over NCCL.
  c5            configs[4]: stream.FindReader over an 8 GiB device-resident stream per GPU (64 GiB on 8), chunk-sharded.
Inputs are larger than L2 (126 MB), so no L2 flush is needed between iterations.  Rank 0 prints ONE JSON line.

  value         device-resident throughput: input already in HBM, results left in HBM (CUDA events on the library's
                stream, max over ranks)
  e2e           the same metric through the host-buffer C-ABI call (pinned host input, H2D + kernels + D2H of the
                results inside the timed region)
  roofline      dominant kernel only: algorithmic bytes / its mean CUDA-event duration vs the measured HBM peak
  parity        after the timed region every rank regenerates a window of ITS input on the CPU, runs the oracle over it
                and compares the device results element-wise; `result_hash` covers the whole device result
  cpu_baseline  the CPU oracle (restated-reference C, NOT Go) on a bounded sample: all host cores, and one core

--impl reference times that CPU oracle as the reference arm (Go is not installed on these boxes, so the reference's
generated Go cannot run; the oracle is the C restatement of the same loops).  That arm loads pattern blobs committed
under tests/golden/ and never loads the product library.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c3": dict(kind="url", pattern_name="URL_PATTERN", blob="c3_url", gib=4.0, api="FindAllBytes",
               desc="URLCapture TDFA (4 named groups) FindAllBytes over 4 GiB synthetic text per GPU (configs[2])"),
    "c2": dict(kind="log", pattern_name="EMAIL_PATTERN", blob="c2_email", gib=1.0, api="FindAllBytes",
               desc="Email (?P<user>\\w+)@(?P<domain>\\w+)\\.(?P<tld>\\w+) FindAllBytes over 1 GiB synthetic log buffer (configs[1])"),
    "c4": dict(kind="suite", api="MatchBytes", inputs=12_500_000,
               desc="corpus pattern suite, batched MatchBytes over 12.5 M short inputs per GPU (100 M on 8 GPUs), one multi-program launch (configs[3])"),
    "c5": dict(kind="stream", pattern_name="DATE_CAPTURE_PATTERN", blob="c5_date", gib=8.0, api="FindReader",
               desc="stream.FindReader, DatePattern, 8 GiB device-resident stream per GPU (64 GiB on 8 GPUs), chunk-sharded (configs[4])"),
}
METRIC = {"FindAllBytes": "GB/s input scanned for FindAllBytes", "MatchBytes": "GB/s input scanned for batched MatchBytes",
          "FindReader": "GB/s input scanned for stream.FindReader"}
BLOB_FILE = os.path.join(ROOT, "tests", "golden", "bench_blobs.npz")


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def committed_blob(name):
    """Program blob committed by tests/golden/make_bench_blobs.py (no product code at run time)."""
    with np.load(BLOB_FILE) as z:
        return z[name].tobytes()


def suite_patterns():
    """The corpus patterns of the batched-MatchBytes workload: unique, ASCII pattern text, no \\p{..}."""
    with open(os.path.join(ROOT, "tests", "golden", "corpus_expected.json")) as fh:
        corpus = json.load(fh)
    seen, out = set(), []
    for ent in corpus["e2e"] + corpus["curated"]:
        pat = ent["pattern"]
        if pat in seen or any(ord(ch) > 127 for ch in pat) or "\\p{" in pat:
            continue
        seen.add(pat)
        out.append((pat, [c["input"] for c in ent["cases"]]))
    return out


def suite_pools(pool_size=2048):
    """Per pattern: a pool of seeded mutations of its corpus inputs (synth.mutate_inputs), packed."""
    from regengo_b200 import synth
    pools = []
    for k, (pat, cases) in enumerate(suite_patterns()):
        pools.append((pat, synth.mutate_inputs(cases, pool_size, stream=k + 1)))
    return pools


def suite_batch(pools, n_inputs, seed_block, device=None):
    """A packed batch of n_inputs short inputs, an equal share per pattern, each drawn from that pattern's pool with a
    counter-based PRNG (seed C4, stream = seed_block).  -> (bytes uint8, offsets int64[n+1], prog_first[n_progs+1])"""
    from regengo_b200 import synth
    rng = synth._rng(synth.SEED_C4, 1_000_000 + seed_block)
    n_p = len(pools)
    per = n_inputs // n_p
    prog_first = np.arange(n_p + 1, dtype=np.uint64) * per
    toks, lens_all = [], []
    base = 0
    idx_parts = []
    for pat, pool in pools:
        toks += pool
        idx_parts.append(base + rng.integers(0, len(pool), size=per))
        base += len(pool)
    pool_obj = synth._Pool([t if len(t) else b"\x00" for t in toks])
    true_len = np.array([len(t) for t in toks], dtype=np.int32)
    pool_obj.lens = true_len          # empty inputs stay empty (the pad byte above is never copied)
    idx = np.concatenate(idx_parts)
    lens = true_len[idx].astype(np.int64)
    offs = np.zeros(idx.size + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    if device is None:
        return pool_obj.concat(idx), offs, prog_first
    import torch
    CH = 1 << 22
    parts = [pool_obj.concat(idx[i:i + CH], device) for i in range(0, idx.size, CH)]
    return torch.cat(parts), offs, prog_first


class CpuOracleSample:
    """A bounded sample of a FindAll / FindReader workload for the CPU oracle: `threads` independent buffers of
    `mib_per_thread` MiB (generated once, in parallel), one oracle instance per thread."""

    def __init__(self, wl, pattern_blob, mib_per_thread, threads, first_block=0):
        from oracle import Oracle
        from regengo_b200 import synth
        self.threads, self.mib, self.api = threads, mib_per_thread, wl["api"]
        self.bufs = [None] * threads
        kw = dict(digit_noise=0.02) if wl["kind"] == "stream" else {}

        def gen(t):
            self.bufs[t] = synth.make_buffer(wl["kind"], mib_per_thread << 20, first_block=first_block + t * mib_per_thread, **kw)
        th = [threading.Thread(target=gen, args=(t,)) for t in range(threads)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        self.oracles = [Oracle(pattern_blob) for _ in range(threads)]
        self.counts = [0] * threads

    def run(self, passes, threads=None):
        """`passes` passes per thread over its buffer; returns wall seconds."""
        threads = threads or self.threads

        def work(t):
            for _ in range(passes):
                if self.api == "FindReader":
                    n = self.oracles[t].find_reader(self.bufs[t], cap=16, count_only=True)[0]
                else:
                    n, _ = self.oracles[t].find_all(self.bufs[t], cap=16, count_only=True)   # one pass, no giant result array
                self.counts[t] = n
        th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t0

    def throughput(self, target_s, max_passes=24, threads=None):
        """One calibration pass, then as many timed passes as fit `target_s` seconds.  -> (GB/s, seconds, passes)"""
        threads = threads or self.threads
        dt1 = self.run(1, threads)
        passes = int(max(1, min(max_passes, round(target_s / max(dt1, 1e-3)))))
        dt = self.run(passes, threads)
        return threads * (self.mib << 20) * passes / dt / 1e9, dt, passes


class CpuSuiteSample:
    """A bounded sample of the batched-MatchBytes workload for the CPU oracle: the same batch generator, one contiguous
    share of the patterns per thread."""

    def __init__(self, n_inputs, threads):
        from oracle import Oracle
        self.threads = threads
        self.pools = suite_pools()
        with np.load(BLOB_FILE) as z:
            self.oracles = [Oracle(z["c4_%03d" % k].tobytes()) for k in range(len(self.pools))]
        self.data, offs, self.first = suite_batch(self.pools, n_inputs, seed_block=7)
        self.offs = offs.astype(np.uint64)
        self.bytes_total = int(self.offs[-1])
        self.n_inputs = int(self.first[-1])

    def run(self, passes, threads=None):
        threads = threads or self.threads
        n_p = len(self.pools)
        shares = [range(t * n_p // threads, (t + 1) * n_p // threads) for t in range(threads)]

        def work(t):
            for _ in range(passes):
                for p in shares[t]:
                    lo, hi = int(self.first[p]), int(self.first[p + 1])
                    self.oracles[p].match_batch(self.data, self.offs[lo:hi + 1])
        th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t0

    def throughput(self, target_s, max_passes=24, threads=None):
        dt1 = self.run(1, threads)
        passes = int(max(1, min(max_passes, round(target_s / max(dt1, 1e-3)))))
        dt = self.run(passes, threads)
        return self.bytes_total * passes / dt / 1e9, dt, passes


def cpu_sample_for(wl, ref_mib, threads):
    if wl["kind"] == "suite":
        return CpuSuiteSample(400_000, threads), "%d threads over the same pattern suite, 400 k inputs of the same generator" % threads
    return (CpuOracleSample(wl, committed_blob(wl["blob"]), ref_mib, threads),
            "%d threads x %d MiB of the same synthetic %s workload" % (threads, ref_mib, wl["kind"]))


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def base_config(args, wl):
    """`config` of the JSON line: what DEFINES the workload, from the command line only -- identical on both arms
    (what a run measured goes to `run_info`)."""
    cfg = {"workload": wl["desc"], "l2": "input larger than L2, no flush needed"}
    if wl["kind"] == "suite":
        cfg["inputs_per_gpu"] = args.inputs if args.inputs else wl["inputs"]
    else:
        cfg["bytes_per_gpu"] = int((args.gib if args.gib is not None else wl["gib"]) * (1 << 30))
    if wl["kind"] == "stream":
        cfg["buffer_size"] = args.buffer_size or 65536
    return cfg


def run_reference(args, wl, rank, world):
    """Reference arm: the CPU oracle (C restatement of the generated loops) on bounded samples."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample_set, sample_desc = cpu_sample_for(wl, args.ref_mib, cores)   # generated once; every step re-times the same sample
    vals = []
    passes = 1
    for s in range(args.warmup + args.steps):
        v, dt, passes = sample_set.throughput(args.ref_seconds)
        if s >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    v1, _, _ = sample_set.throughput(min(args.ref_seconds, 3.0), threads=1)
    sample = f"{sample_desc} x {passes} passes per step, {wl['api']}, restated-reference CPU baseline (C oracle), not Go"
    line = {
        "impl": "reference", "metric": METRIC[wl["api"]], "value": value, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": base_config(args, wl),
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample, "one_core_value": v1},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def tensor_hash(*tensors):
    """Order-sensitive 64-bit hash of device tensors (wrapping int64 arithmetic on the device)."""
    import torch
    h = 0
    for t in tensors:
        v = t.reshape(-1).to(torch.int64)
        idx = torch.arange(1, v.numel() + 1, device=v.device, dtype=torch.int64)
        h = (h * 1000003 + int(((v + 0x9E3779B9) * (idx * 2654435761 + 1)).sum().item())) & 0xFFFFFFFFFFFFFFFF
    return h


def rle_records(recs):
    """Expanded FindAll list (the reference's) -> run-length form (records, reps), the device's result layout."""
    if recs.shape[0] == 0:
        return recs, np.zeros(0, dtype=np.int64)
    change = np.ones(recs.shape[0], dtype=bool)
    change[1:] = (recs[1:] != recs[:-1]).any(axis=1)
    starts = np.nonzero(change)[0]
    reps = np.diff(np.append(starts, recs.shape[0]))
    return recs[starts], reps


# ---------------------------------------------------------------------------------------------------------------------
# workloads on the device
# ---------------------------------------------------------------------------------------------------------------------
class FindAllWork:
    """c2 / c3: FindAllBytes over one large buffer (N > 1: one logical buffer, sharded)."""

    def __init__(self, args, wl, env):
        import torch
        import regengo_b200 as rg
        from regengo_b200 import _lib, synth
        from regengo_b200 import dist as rdist
        self.args, self.wl, self.env = args, wl, env
        self.torch, self.rg, self.synth, self.rdist = torch, rg, synth, rdist
        self.L = _lib.load()
        self._lib = _lib
        rank, world, dev = env["rank"], env["world"], env["dev"]
        self.pat = rg.Pattern(getattr(synth, wl["pattern_name"]), device=env["local_rank"])
        self.ctx = rg.context(env["local_rank"])
        _lib.check(self.L.rgx_ctx_enable_timing(self.ctx, 1))
        gib = args.gib if args.gib is not None else wl["gib"]
        self.n_bytes = int(gib * (1 << 30))
        self.blocks = (self.n_bytes + synth.BLOCK - 1) // synth.BLOCK
        t0 = time.perf_counter()
        # N > 1: ONE logical buffer of world * n_bytes, rank r holds bytes [r*n_bytes, (r+1)*n_bytes) plus a halo (the
        # next rank's first MiB) and a pre-halo (the previous rank's last MiB), both regenerated locally from the same
        # block generator: no exchange of input bytes
        self.halo = synth.BLOCK if (world > 1 and rank < world - 1) else 0
        self.pre = synth.BLOCK if (world > 1 and rank > 0) else 0
        self.buf = synth.make_buffer(wl["kind"], self.pre + self.n_bytes + self.halo, first_block=rank * self.blocks - (1 if self.pre else 0), device=dev)
        torch.cuda.synchronize()
        self.gen_s = time.perf_counter() - t0
        self.nc = self.pat.num_cap
        self.cap_rec = self.n_bytes // 64 + 1024
        self.d_out = torch.empty(self.cap_rec * self.nc, dtype=torch.int64, device=dev)
        self.d_reps = torch.empty(self.cap_rec, dtype=torch.int32, device=dev)
        self.n_rec = C.c_uint64()
        self.exit_cur = C.c_int64()
        self.shard_start = rank * self.n_bytes
        self.gather_pair = rdist.torch_all_gather_pair(device=dev) if world > 1 else None
        self.gather_post, self.gather_collect = rdist.torch_all_gather_pair_async(device=dev) if world > 1 else (None, None)
        self.pending = None
        self.redos = 0
        self.exchange_rounds = 0
        self.entry_global = 0
        self.phase = (C.c_float * 4)()
        self.step_phase = [0.0, 0.0, 0.0]
        self.phases = []
        self.total_matches = 0
        plan = self.pat.device_plan()
        self.kernel = ("findall_scan6_kernel" if plan.get("fast_tdfa_scan") else "findall_scan_btrun_kernel" if plan.get("run_anchor")
                       else "findall_scan_linear_kernel" if plan.get("linear_findall_scan") else "findall_scan_kernel")
        self.algorithmic_bytes = self.n_bytes

    def step(self):
        L, ctx, pat, world, rank = self.L, self.ctx, self.pat, self.env["world"], self.env["rank"]
        if world == 1:
            r = L.rgx_find_all_dev(ctx, pat._h, self.buf.data_ptr(), self.n_bytes, -1, self.d_out.data_ptr(), self.d_reps.data_ptr(),
                                   self.cap_rec, C.byref(self.n_rec))
            self._lib.check(r)
            L.rgx_ctx_last_timing(ctx, self.phase)
            self.step_phase = [self.phase[0], self.phase[1], self.phase[2]]
            self.total_matches = r
            return r
        # sharded: every rank > 0 carries the cursor through its pre-halo, ONE 16-byte all-gather confirms entry == the
        # predecessor's exit (a rank whose carried cursor was wrong redoes its shard from the right entry).  The
        # confirmation of step k is posted when its kernels are done and collected while step k + 1 computes; drain()
        # collects the last one inside the timed region.
        entry, exit_cur, total_local = self._run_pre(self.buf.data_ptr())
        handle = self.gather_post((entry, exit_cur))
        self._confirm_pending()
        self.pending = handle
        self.exchange_rounds = 1
        self.entry_global = entry
        self.total_matches = total_local
        return total_local

    def _confirm_pending(self):
        if self.pending is None:
            return
        pairs = self.gather_collect(self.pending)
        self.pending = None
        if self.rdist.pre_halo_bad_ranks(pairs, self.env["world"]):
            # (never seen: the pre-halo holds thousands of matches)  that step again, synchronously, with the redo protocol
            self.redos += 1
            self._sharded_pass(self.buf.data_ptr())

    def drain(self):
        if self.env["world"] > 1:
            self._confirm_pending()

    def _run_pre(self, buf_ptr):
        """Scan + replay + output of this rank's shard with its pre-halo -> (entry_global, exit_global, matches)."""
        L, ctx, pat, world, rank = self.L, self.ctx, self.pat, self.env["world"], self.env["rank"]
        is_last = int(rank == world - 1)
        entry_rel, exit_rel = C.c_int64(), C.c_int64()
        if rank == 0:
            r = L.rgx_find_all_shard_dev(ctx, pat._h, buf_ptr, self.n_bytes + self.halo, self.n_bytes, is_last, 0, 0, 0,
                                         self.d_out.data_ptr(), self.d_reps.data_ptr(), self.cap_rec, C.byref(self.n_rec), C.byref(exit_rel))
            self._lib.check(r)
            entry, exit_cur = 0, exit_rel.value
        else:
            r = L.rgx_find_all_shard_pre_dev(ctx, pat._h, buf_ptr, self.pre + self.n_bytes + self.halo, self.pre, self.n_bytes, is_last,
                                             self.shard_start, self.d_out.data_ptr(), self.d_reps.data_ptr(), self.cap_rec,
                                             C.byref(self.n_rec), C.byref(entry_rel), C.byref(exit_rel))
            self._lib.check(r)
            entry, exit_cur = self.shard_start + entry_rel.value, self.shard_start + exit_rel.value
        L.rgx_ctx_last_timing(ctx, self.phase)
        self.step_phase = [self.phase[0], self.phase[1], self.phase[2]]
        return entry, exit_cur, r

    def _sharded_pass(self, buf_ptr):
        L, ctx, pat, world, rank = self.L, self.ctx, self.pat, self.env["world"], self.env["rank"]
        is_last = int(rank == world - 1)
        entry_rel, exit_rel = C.c_int64(), C.c_int64()

        def timing():
            L.rgx_ctx_last_timing(ctx, self.phase)
            self.step_phase = [self.phase[0], self.phase[1], self.phase[2]]

        def run_pre():
            if rank == 0:
                r = L.rgx_find_all_shard_dev(ctx, pat._h, buf_ptr, self.n_bytes + self.halo, self.n_bytes, is_last, 0, 0, 0,
                                             self.d_out.data_ptr(), self.d_reps.data_ptr(), self.cap_rec, C.byref(self.n_rec), C.byref(exit_rel))
                self._lib.check(r)
                timing()
                return 0, exit_rel.value, r
            r = L.rgx_find_all_shard_pre_dev(ctx, pat._h, buf_ptr, self.pre + self.n_bytes + self.halo, self.pre, self.n_bytes, is_last,
                                             self.shard_start, self.d_out.data_ptr(), self.d_reps.data_ptr(), self.cap_rec,
                                             C.byref(self.n_rec), C.byref(entry_rel), C.byref(exit_rel))
            self._lib.check(r)
            timing()
            return self.shard_start + entry_rel.value, self.shard_start + exit_rel.value, r

        def run_from(entry_global):
            self.redos += 1
            r = L.rgx_find_all_shard_dev(ctx, pat._h, buf_ptr + self.pre, self.n_bytes + self.halo, self.n_bytes, is_last,
                                         entry_global - self.shard_start, self.shard_start, 0, self.d_out.data_ptr(), self.d_reps.data_ptr(),
                                         self.cap_rec, C.byref(self.n_rec), C.byref(exit_rel))
            self._lib.check(r)
            return self.shard_start + exit_rel.value, r
        entry, _, total_local, rounds = self.rdist.settle_pre_halo_chain(run_pre, run_from, rank, world, self.shard_start, self.gather_pair)
        return total_local, entry, rounds

    def after_step(self):
        self.phases.append(tuple(self.step_phase))

    def result_tensors(self):
        n = int(self.n_rec.value)
        return [self.d_out[: n * self.nc], self.d_reps[:n]]

    def e2e_sharded(self, k_e2e, barrier):
        """N > 1: the SAME sharded computation from host buffers -- every step uploads the rank's part of the logical
        buffer (shard + halos) from pinned memory, runs the sharded pass (with its cursor all-gather) and downloads the
        rank's records."""
        torch = self.torch
        held = self.pre + self.n_bytes + self.halo
        h_in = torch.empty(held, dtype=torch.uint8, pin_memory=True)
        h_in.copy_(self.buf)
        d_in = torch.empty(held + 16, dtype=torch.uint8, device=self.env["dev"])
        n_rec_dev = int(self.n_rec.value)
        cap_e2e = min(self.cap_rec, n_rec_dev + n_rec_dev // 8 + 65536)
        h_out = torch.empty(cap_e2e * self.nc, dtype=torch.int64, pin_memory=True)
        h_reps = torch.empty(cap_e2e, dtype=torch.int32, pin_memory=True)
        stream = self.env["stream"]
        torch.cuda.synchronize()
        tot = [0]

        def e2e_step():
            with torch.cuda.stream(stream):
                d_in[:held].copy_(h_in, non_blocking=True)
            tot[0], _, _ = self._sharded_pass(d_in.data_ptr())
            n = int(self.n_rec.value)
            with torch.cuda.stream(stream):
                h_out[: n * self.nc].copy_(self.d_out[: n * self.nc], non_blocking=True)
                h_reps[:n].copy_(self.d_reps[:n], non_blocking=True)
            stream.synchronize()
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / k_e2e
        assert tot[0] == self.total_matches and int(self.n_rec.value) == n_rec_dev
        return dt, held, n_rec_dev * (self.nc * 8 + 4) + 16, "rgx_find_all_shard_pre_dev on the rank's part of the logical buffer; pinned H2D / D2H copies and the cursor all-gather inside the step"

    def e2e(self, k_e2e, barrier):
        if self.env["world"] > 1:
            return self.e2e_sharded(k_e2e, barrier)
        torch, L = self.torch, self.L
        h_in = torch.empty(self.n_bytes, dtype=torch.uint8, pin_memory=True)
        h_in.copy_(self.buf[:self.n_bytes])
        # result capacity: what the device-resident steps produced plus slack; keeps the pinned allocation near 1 GB
        n_rec_dev = int(self.n_rec.value)
        cap_e2e = min(self.cap_rec, n_rec_dev + n_rec_dev // 8 + 65536)
        h_out = torch.empty(cap_e2e * self.nc, dtype=torch.int64, pin_memory=True)
        h_reps = torch.empty(cap_e2e, dtype=torch.int32, pin_memory=True)
        torch.cuda.synchronize()
        n_rec = C.c_uint64()

        def e2e_step():
            return self._lib.check(L.rgx_find_all_rle(self.ctx, self.pat._h, h_in.data_ptr(), self.n_bytes, -1, h_out.data_ptr(),
                                                      h_reps.data_ptr(), cap_e2e, C.byref(n_rec)))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            tot = e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / k_e2e
        if self.env["world"] == 1:
            assert tot == self.total_matches and n_rec.value == n_rec_dev
        return dt, self.n_bytes, int(n_rec.value) * (self.nc * 8 + 4) + 256, "rgx_find_all_rle (host buffers, run-length result records)"

    def parity(self, window_mib):
        """The rank's records from its entry cursor to the end of a window at the start of its shard, against the CPU
        oracle run over the same bytes regenerated on the host (synth is block-addressable)."""
        from oracle import Oracle
        synth = self.synth
        win = min(window_mib << 20, self.n_bytes)
        cur = self.entry_global                      # global position of the cursor entering this shard (rank 0: 0)
        blk0 = cur // synth.BLOCK
        gen_start = blk0 * synth.BLOCK
        want_end = self.shard_start + win
        tail = synth.BLOCK                            # bytes past the window so that matches near its end are whole
        total_logical = self.env["world"] * self.n_bytes
        gen_end = min(want_end + tail, total_logical)
        host = synth.make_buffer(self.wl["kind"], gen_end - gen_start, first_block=blk0)
        o = Oracle(self.pat.blob())
        cnt, recs = o.find_all(host[cur - gen_start:])
        recs = recs + cur                             # slice-relative -> global offsets (nil groups stay -1 below)
        recs[recs < cur] = -1
        erecs, ereps = rle_records(recs)
        keep = erecs[:, 0] < want_end
        erecs, ereps = erecs[keep], ereps[keep]
        n = int(self.n_rec.value)
        d_recs = self.d_out[: n * self.nc].view(-1, self.nc)
        m = int((d_recs[:, 0] < want_end).sum().item())
        got = d_recs[:m].cpu().numpy()
        got_reps = self.d_reps[:m].cpu().numpy().astype(np.int64)
        ok = bool(got.shape == erecs.shape and np.array_equal(got, erecs) and np.array_equal(got_reps, ereps))
        return {"window_bytes": int(want_end - cur), "records": int(erecs.shape[0]), "matches": int(ereps.sum()), "ok": ok,
                "checked": "records and repeat counts, element-wise, CPU oracle over the regenerated window"}

    def config_extra(self):
        world = self.env["world"]
        return {"bytes_per_gpu": self.n_bytes, "matches_per_step": int(self.total_matches), "distinct_records_per_step": int(self.n_rec.value),
                "result_form": "run-length offset records left in HBM", "gen_seconds": self.gen_s,
                "sharding": None if world == 1 else f"one logical buffer of {world}x{self.n_bytes} B, 1 MiB halo + 1 MiB pre-halo per rank, one "
                            f"16-byte all_gather (NCCL) per step confirms the carried cursors (posted when the step's kernels are done, collected while the next step computes; the last one inside the timed region): {self.redos} redo(s) on rank 0"}

    def units(self):
        return self.n_bytes

    def roofline_extra(self):
        ph = self.phases
        return {"scan_ms": float(np.mean([s[0] for s in ph])), "chain_ms": float(np.mean([s[1] for s in ph])),
                "emit_ms": float(np.mean([s[2] for s in ph]))}

    def dominant_ms(self):
        return float(np.mean([s[0] for s in self.phases]))


class SuiteWork:
    """c4: batched MatchBytes of the corpus suite, one multi-program launch, inputs sharded over the ranks."""

    def __init__(self, args, wl, env):
        import torch
        import regengo_b200 as rg
        from regengo_b200 import _lib
        from regengo_b200 import dist as rdist
        self.args, self.wl, self.env, self.torch, self.rg, self._lib, self.rdist = args, wl, env, torch, rg, _lib, rdist
        self.L = _lib.load()
        dev = env["dev"]
        self.ctx = rg.context(env["local_rank"])
        t0 = time.perf_counter()
        self.pools = suite_pools()
        self.pats = [rg.Pattern(p, device=env["local_rank"]) for p, _ in self.pools]
        n_inputs = args.inputs if args.inputs else wl["inputs"]
        self.data, offs, self.first = suite_batch(self.pools, n_inputs, seed_block=env["rank"], device=dev)
        self.n_inputs = int(self.first[-1])
        self.n_bytes = int(offs[-1])
        assert self.n_bytes < (1 << 32)
        self.h_offs = offs
        self.d_offs = torch.from_numpy(offs.astype(np.uint32).view(np.int32)).to(dev)
        self.d_out = torch.empty(self.n_inputs, dtype=torch.uint8, device=dev)
        self.handles = (C.c_void_p * len(self.pats))(*[p._h for p in self.pats])
        self.pf = np.ascontiguousarray(self.first, dtype=np.uint64)
        torch.cuda.synchronize()
        self.gen_s = time.perf_counter() - t0
        self.kernel = "match_multi_kernel"
        self.algorithmic_bytes = self.n_bytes + 4 * self.n_inputs + self.n_inputs
        self.gathered = None
        self.ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        self.stream = env["stream"]
        self.kernel_ms = []

    def step(self):
        self.ev[0].record(self.stream)
        self._lib.check(self.L.rgx_match_multi_dev(self.ctx, self.handles, len(self.pats), self.data.data_ptr(), self.d_offs.data_ptr(),
                                                   self.pf.ctypes.data, self.d_out.data_ptr()))
        self.ev[1].record(self.stream)
        if self.env["world"] > 1:
            # the path's one collective: gather the 1-byte flags of every rank on rank 0 (NCCL)
            self.torch.cuda.current_stream().wait_stream(self.stream)
            self.gathered = self.rdist.gather_flags(self.d_out)
        return self.n_inputs

    def after_step(self):
        self.torch.cuda.synchronize()
        self.kernel_ms.append(self.ev[0].elapsed_time(self.ev[1]))

    def result_tensors(self):
        return [self.d_out]

    def e2e(self, k_e2e, barrier):
        torch, L = self.torch, self.L
        h_bytes = torch.empty(self.n_bytes, dtype=torch.uint8, pin_memory=True)
        h_bytes.copy_(self.data)
        h_offs = torch.from_numpy(self.h_offs.astype(np.uint64).view(np.int64)).pin_memory()
        h_out = torch.empty(self.n_inputs, dtype=torch.uint8, pin_memory=True)
        torch.cuda.synchronize()

        def e2e_step():
            return self._lib.check(L.rgx_match_multi(self.ctx, self.handles, len(self.pats), h_bytes.data_ptr(), h_offs.data_ptr(),
                                                     self.pf.ctypes.data, h_out.data_ptr()))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / k_e2e
        assert bool((h_out == self.d_out.cpu()).all())
        return dt, self.n_bytes + 8 * (self.n_inputs + 1), self.n_inputs, "rgx_match_multi (host buffers: bytes and 64-bit offsets uploaded, flags downloaded)"

    def parity(self, window_mib):
        """Every pattern's first inputs (about a million in all) against the CPU oracle's MatchBytes."""
        from oracle import Oracle
        per = max(1, min(1_000_000 // len(self.pats), int(self.first[1] - self.first[0])))
        host_flags = self.d_out.cpu().numpy()
        checked, bad = 0, 0
        for k, p in enumerate(self.pats):
            lo = int(self.first[k]); hi = lo + per
            b0, b1 = int(self.h_offs[lo]), int(self.h_offs[hi])
            data = self.data[b0:b1].cpu().numpy()
            offs = (self.h_offs[lo:hi + 1] - b0).astype(np.uint64)
            exp = Oracle(p.blob()).match_batch(data, offs)
            bad += int((exp != host_flags[lo:hi]).sum())
            checked += per
        return {"window_inputs": checked, "mismatches": bad, "ok": bad == 0,
                "checked": "MatchBytes flags of the first inputs of every pattern, CPU oracle on the same bytes"}

    def config_extra(self):
        world = self.env["world"]
        return {"patterns": len(self.pats), "inputs_per_gpu": self.n_inputs, "bytes_per_gpu": self.n_bytes,
                "mean_input_bytes": self.n_bytes / self.n_inputs, "inputs_per_s": None,
                "inputs": "per pattern a pool of 2048 seeded mutations of its corpus inputs, drawn with a counter-based PRNG",
                "matched_fraction": float(self.d_out.float().mean().item()), "gen_seconds": self.gen_s,
                "sharding": None if world == 1 else f"inputs sharded over {world} ranks, 1-byte flags gathered on rank 0 (NCCL all_gather) inside the step"}

    def units(self):
        return self.n_bytes

    def roofline_extra(self):
        return {"kernel_ms": float(np.mean(self.kernel_ms)), "inputs_per_s": self.n_inputs / (float(np.mean(self.kernel_ms)) * 1e-3),
                "algorithmic_bytes_note": "input bytes + 4 B offset + 1 B flag per input"}

    def dominant_ms(self):
        return float(np.mean(self.kernel_ms))


class ReaderWork:
    """c5: stream.FindReader over a device-resident stream, chunk ranges sharded over the ranks."""

    def __init__(self, args, wl, env):
        import torch
        import regengo_b200 as rg
        from regengo_b200 import _lib, synth
        from regengo_b200 import dist as rdist
        self.args, self.wl, self.env, self.torch, self.rg, self._lib, self.synth = args, wl, env, torch, rg, _lib, synth
        self.L = _lib.load()
        rank, world, dev = env["rank"], env["world"], env["dev"]
        self.pat = rg.Pattern(getattr(synth, wl["pattern_name"]), device=env["local_rank"])
        self.ctx = rg.context(env["local_rank"])
        gib = args.gib if args.gib is not None else wl["gib"]
        per_rank = int(gib * (1 << 30))
        self.total_len = per_rank * world
        cfg = self.pat.stream_config(rg.StreamConfig(args.buffer_size, 0))
        self.B, self.Lo = cfg.buffer_size, cfg.max_leftover
        stride = self.B - self.Lo
        # the chunk schedule of a reader that fills every Read (streaming.go:123-245; capi_stream.inc make_chunk_plan)
        full = (self.total_len - self.B) // stride + 1 if self.total_len >= self.B else 0
        exact = full > 0 and (self.total_len - self.B) % stride == 0
        n_chunks_total = (full + 1 if self.Lo > 0 else full) if exact else full + 1
        self.first_chunk, self.n_chunks = rdist.shard_chunks(n_chunks_total, world, rank)
        lo, hi = rdist.chunk_span(self.first_chunk, self.n_chunks, self.B, self.Lo, self.total_len)
        self.base, self.n_bytes = lo, hi - lo
        t0 = time.perf_counter()
        blk0 = lo // synth.BLOCK
        pre = lo - blk0 * synth.BLOCK
        full = synth.make_buffer("stream", pre + self.n_bytes, first_block=blk0, device=dev, digit_noise=0.02)
        self.buf = full[pre:]
        torch.cuda.synchronize()
        self.gen_s = time.perf_counter() - t0
        self.nc = self.pat.num_cap
        self.cap = self.n_bytes // 40 + 1024
        self.d_so = torch.empty(self.cap, dtype=torch.int64, device=dev)
        self.d_ci = torch.empty(self.cap, dtype=torch.int32, device=dev)
        self.d_rec = torch.empty(self.cap * self.nc, dtype=torch.int64, device=dev)
        self.count = 0
        self.kernel = "find_reader kernels"
        self.algorithmic_bytes = self.n_bytes
        self.ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        self.stream = env["stream"]
        self.kernel_ms = []

    def step(self):
        self.ev[0].record(self.stream)
        self.count = self._lib.check(self.L.rgx_find_reader_dev(self.ctx, self.pat._h, self.buf.data_ptr(), self.base, self.n_bytes, self.total_len,
                                                                self.args.buffer_size, 0, self.first_chunk, self.n_chunks,
                                                                self.d_so.data_ptr(), self.d_ci.data_ptr(), self.d_rec.data_ptr(), self.cap))
        self.ev[1].record(self.stream)
        return self.count

    def after_step(self):
        self.torch.cuda.synchronize()
        self.kernel_ms.append(self.ev[0].elapsed_time(self.ev[1]))

    def result_tensors(self):
        n = int(self.count)
        return [self.d_so[:n], self.d_ci[:n], self.d_rec[: n * self.nc]]

    def e2e(self, k_e2e, barrier):
        torch, L = self.torch, self.L
        # a rank-local stream through the host-buffer entry point (the whole shard as one reader)
        n = min(self.n_bytes, 2 << 30)
        h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        h_in.copy_(self.buf[:n])
        cap = n // 40 + 1024
        so = torch.empty(cap, dtype=torch.int64, pin_memory=True)
        ci = torch.empty(cap, dtype=torch.int32, pin_memory=True)
        rec = torch.empty(cap * self.nc, dtype=torch.int64, pin_memory=True)
        torch.cuda.synchronize()

        def e2e_step():
            return self._lib.check(L.rgx_find_reader(self.ctx, self.pat._h, h_in.data_ptr(), n, self.args.buffer_size, 0, 0, -1,
                                                     so.data_ptr(), ci.data_ptr(), rec.data_ptr(), cap))
        cnt = e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / k_e2e
        # scale: the e2e sample is `n` bytes of the rank's stream
        self.e2e_bytes = n
        return dt, n, int(cnt) * (self.nc * 8 + 12), "rgx_find_reader (host buffers; the first 2 GiB of the rank's stream)"

    def parity(self, window_mib):
        """The first chunks of the rank's range against the CPU oracle's FindReader over the same bytes (chunks are
        independent given BufferSize / MaxLeftover, so a reader started at a chunk boundary replays them)."""
        from oracle import Oracle
        stride = self.B - self.Lo
        n_win = max(3, min(self.n_chunks, (window_mib << 20) // stride))
        nbytes = min(self.n_bytes, n_win * stride + self.B)
        host = self.buf[:nbytes].cpu().numpy()
        cnt, so, ci, recs = Oracle(self.pat.blob()).find_reader(host, self.args.buffer_size, 0)
        last_full = n_win - 2 if nbytes < self.n_bytes else n_win          # the oracle's final (short) chunks differ from mid-stream ones
        keep = ci < last_full
        so, ci, recs = so[keep] + self.base, ci[keep] + self.first_chunk, recs[keep] + self.base
        n = int(self.count)
        d_ci = self.d_ci[:n]
        m = int((d_ci < self.first_chunk + last_full).sum().item())
        ok = bool(m == so.shape[0] and np.array_equal(self.d_so[:m].cpu().numpy(), so) and np.array_equal(d_ci[:m].cpu().numpy(), ci)
                  and np.array_equal(self.d_rec[: m * self.nc].view(-1, self.nc).cpu().numpy(), recs))
        return {"window_bytes": int(last_full * stride), "records": int(so.shape[0]), "ok": ok,
                "checked": "(StreamOffset, ChunkIndex, offsets) of every match of the window's chunks, CPU oracle over the same bytes"}

    def config_extra(self):
        world = self.env["world"]
        return {"bytes_per_gpu": self.n_bytes, "buffer_size": self.B, "max_leftover": self.Lo, "chunks_per_gpu": self.n_chunks,
                "matches_per_step": int(self.count), "digit_noise": 0.02, "gen_seconds": self.gen_s,
                "sharding": None if world == 1 else f"chunk ranges over {world} ranks, each holding its span plus the {self.Lo}-byte halo; no data-path collective"}

    def units(self):
        return self.n_bytes

    def roofline_extra(self):
        return {"kernel_ms": float(np.mean(self.kernel_ms))}

    def dominant_ms(self):
        return float(np.mean(self.kernel_ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--gib", type=float, default=None, help="GiB of input per GPU (default: the config's size)")
    ap.add_argument("--inputs", type=int, default=None, help="c4: inputs per GPU (default 12.5 M)")
    ap.add_argument("--buffer-size", type=int, default=0, help="c5: stream.Config.BufferSize (0 = the pattern's default, 64 KiB)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-mib", type=int, default=32, help="MiB per host thread per step for the CPU arm")
    ap.add_argument("--ref-seconds", type=float, default=4.0, help="wall seconds of CPU work per step of the reference arm")
    ap.add_argument("--parity-mib", type=int, default=64, help="window of the in-run oracle check")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import torch
    import torch.distributed as dist
    import regengo_b200 as rg
    from regengo_b200 import _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.load()
    ctx = rg.context(local_rank)
    stream = torch.cuda.ExternalStream(L.rgx_ctx_stream(ctx), device=dev)
    env = {"rank": rank, "world": world, "local_rank": local_rank, "dev": dev, "stream": stream}
    work = {"FindAllBytes": FindAllWork, "MatchBytes": SuiteWork, "FindReader": ReaderWork}[wl["api"]](args, wl, env)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        work.step()
    if hasattr(work, "drain"):
        work.drain()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = rg.launches(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    work.phases = [] if hasattr(work, "phases") else None
    if hasattr(work, "kernel_ms"):
        work.kernel_ms = []
    t_wall0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        work.step()
        work.after_step()
    if hasattr(work, "drain"):
        work.drain()          # (sharded FindAll: the last step's cursor confirmation is part of the timed region)
    ev1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    sampler.stop_flag = True
    ms_total = ev0.elapsed_time(ev1)
    if wl["api"] != "FindAllBytes":
        ms_total = max(ms_total, wall_ms) if world > 1 else ms_total   # (the gather runs on torch's stream: count it)
    launches = rg.launches(local_rank) - launches0
    total_matches = int(getattr(work, "total_matches", 0) or getattr(work, "count", 0) or 0)
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * work.units() / (ms_per_step * 1e-3) / 1e9

    # result hash + in-run parity (every rank checks its own shard; rank 0 reports the conjunction)
    parity = None
    if not args.no_parity:
        par = work.parity(args.parity_mib)
        h = tensor_hash(*work.result_tensors())
        if world > 1:
            ok_t = torch.tensor([int(par["ok"])], device=dev)
            dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
            hs = [None] * world
            dist.all_gather_object(hs, "%016x" % h)
            par["ok"] = bool(ok_t.item())
            par["ranks_checked"] = world
            par["result_hash"] = hs
        else:
            par["result_hash"] = "%016x" % h
        parity = par

    # end to end through the host-buffer C ABI (rank-local; max over ranks)
    e2e = None
    if not args.no_e2e:
        k_e2e = max(3, min(args.steps, 5))
        dt, h2d, d2h, api = work.e2e(k_e2e, barrier)
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e_units = getattr(work, "e2e_bytes", work.units())
        e2e = {"value": world * e2e_units / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": dt * 1e3, "api": api}

    if rank == 0:
        peak, peak_src = measured_peak()
        dom_ms = work.dominant_ms()
        achieved = work.algorithmic_bytes / (dom_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                tj = json.load(fh)
                traffic, traffic_src = tj.get(args.workload), tj.get("source")
        except Exception:
            pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_source": traffic_src, "kernel": work.kernel, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": work.algorithmic_bytes, "whole_step_frac": value / world / peak}
        roofline.update(work.roofline_extra())
        cpu = None
        if not args.no_cpu:
            cores = os.cpu_count() or 1
            sample_set, sample_desc = cpu_sample_for(wl, args.ref_mib, cores)
            v, dt, passes = sample_set.throughput(8.0)
            v1, dt1, p1 = sample_set.throughput(3.0, threads=1)
            cpu = {"value": v, "unit": "GB/s", "cores": cores, "kind": "port",
                   "sample": f"{sample_desc} x {passes} passes, {wl['api']}, C oracle (restated reference, not Go)",
                   "seconds": dt, "one_core_value": v1}
        cfg = base_config(args, wl)
        info = work.config_extra()
        if "inputs_per_s" in info:
            info["inputs_per_s"] = world * info["inputs_per_gpu"] / (ms_per_step * 1e-3)
        line = {
            "metric": METRIC[wl["api"]], "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": cfg, "run_info": info, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "parity": parity,
            "gpu_launches": int(launches), "clocks": sampler.result(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
