#!/usr/bin/env python3
"""bench.py -- FindAllBytes throughput (GB/s of input scanned) on synthetic buffers, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2] [--gib G] [--impl reference]

A "step" is one FindAllBytes pass of the device path over one resident buffer (default workload c3:
the curated URLCapture TDFA pattern over 4 GiB of synthetic prose per GPU, BASELINE.json configs[2],
the configuration the north_star's target is quoted on; --workload c2 is configs[1], the email
backtracking pattern over 1 GiB).  Inputs are larger than L2 (126 MB), so no L2 flush is needed
between iterations.  One JSON line is printed by rank 0.

  value     device-resident throughput: input already in HBM, results left in HBM (CUDA events on the
            library's stream, max over ranks)
  e2e       the same metric through the host-buffer C-ABI call rgx_find_all_rle (pinned host input,
            H2D + kernels + D2H of the results inside the timed region)
  roofline  scan kernel only: input bytes / its mean CUDA-event duration vs the measured HBM peak
  cpu_baseline  the CPU oracle (restated-reference C, NOT Go) on a bounded sample, all host cores

--impl reference times that CPU oracle as the reference arm (Go is not installed on these boxes, so
the reference's generated Go cannot run; the oracle is the C restatement of the same loops).
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
    "c3": dict(kind="url", pattern_name="URL_PATTERN", gib=4.0,
               desc="URLCapture TDFA (4 named groups) FindAllBytes over 4 GiB synthetic text per GPU (configs[2])"),
    "c2": dict(kind="log", pattern_name="EMAIL_PATTERN", gib=1.0,
               desc="Email (?P<user>\\w+)@(?P<domain>\\w+)\\.(?P<tld>\\w+) FindAllBytes over 1 GiB synthetic log buffer (configs[1])"),
}
METRIC = "GB/s input scanned for FindAllBytes"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class CpuOracleSample:
    """A bounded sample of a workload for the CPU oracle: `threads` independent buffers of `mib_per_thread`
    MiB (generated once, in parallel), one oracle instance per thread."""

    def __init__(self, kind, pattern_blob, mib_per_thread, threads, first_block=0):
        from oracle import Oracle
        from regengo_b200 import synth
        self.threads, self.mib = threads, mib_per_thread
        self.bufs = [None] * threads

        def gen(t):
            self.bufs[t] = synth.make_buffer(kind, mib_per_thread << 20, first_block=first_block + t * mib_per_thread)
        th = [threading.Thread(target=gen, args=(t,)) for t in range(threads)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        self.oracles = [Oracle(pattern_blob) for _ in range(threads)]
        self.counts = [0] * threads

    def run(self, passes):
        """`passes` FindAllBytes(-1) passes per thread over its buffer; returns wall seconds."""
        def work(t):
            for _ in range(passes):
                n, _ = self.oracles[t].find_all(self.bufs[t], cap=16)   # cap: count only, no giant result array
                self.counts[t] = n
        th = [threading.Thread(target=work, args=(t,)) for t in range(self.threads)]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t0

    def throughput(self, target_s, max_passes=24):
        """One calibration pass, then as many timed passes as fit `target_s` seconds.  -> (GB/s, seconds, passes)"""
        dt1 = self.run(1)
        passes = int(max(1, min(max_passes, round(target_s / max(dt1, 1e-3)))))
        dt = self.run(passes)
        return self.threads * (self.mib << 20) * passes / dt / 1e9, dt, passes


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


def run_reference(args, wl, rank, world):
    """Reference arm: the CPU oracle (C restatement of the generated loops) on bounded samples."""
    if rank != 0:
        return
    import regengo_b200 as rg
    from regengo_b200 import synth
    blob = rg.Pattern(getattr(synth, wl["pattern_name"])).blob()   # front-end only: works without a GPU
    cores = os.cpu_count() or 1
    mib = args.ref_mib
    sample_set = CpuOracleSample(wl["kind"], blob, mib, cores)   # generated once; every step re-times the same sample
    vals = []
    passes = 1
    for s in range(args.warmup + args.steps):
        v, dt, passes = sample_set.throughput(args.ref_seconds)
        if s >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    sample = f"{cores} threads x {mib} MiB of the same synthetic {wl['kind']} workload x {passes} passes per step, FindAllBytes(-1)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": wl["desc"], "note": "restated-reference CPU baseline (C oracle), not Go: no Go toolchain on the box"},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--gib", type=float, default=None, help="GiB of input per GPU (default: the config's size)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-mib", type=int, default=32, help="MiB per host thread per step for the CPU arm")
    ap.add_argument("--ref-seconds", type=float, default=4.0, help="wall seconds of CPU work per step of the reference arm")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
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
    from regengo_b200 import _lib, synth
    from regengo_b200 import dist as rdist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.load()
    pat = rg.Pattern(getattr(synth, wl["pattern_name"]), device=local_rank)
    ctx = rg.context(local_rank)
    _lib.check(L.rgx_ctx_enable_timing(ctx, 1))
    stream = torch.cuda.ExternalStream(L.rgx_ctx_stream(ctx), device=dev)

    gib = args.gib if args.gib is not None else wl["gib"]
    n_bytes = int(gib * (1 << 30))
    blocks = (n_bytes + synth.BLOCK - 1) // synth.BLOCK
    t0 = time.perf_counter()
    # N > 1: ONE logical buffer of world * n_bytes, rank r holds bytes [r*n_bytes, (r+1)*n_bytes) plus a
    # halo (the next rank's first MiB, regenerated locally from the same block generator: no exchange)
    halo = synth.BLOCK if (world > 1 and rank < world - 1) else 0
    buf = synth.make_buffer(wl["kind"], n_bytes + halo, first_block=rank * blocks, device=dev)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0

    nc = pat.num_cap
    cap_rec = n_bytes // 64 + 1024
    d_out = torch.empty(cap_rec * nc, dtype=torch.int64, device=dev)
    d_reps = torch.empty(cap_rec, dtype=torch.int32, device=dev)
    n_rec = C.c_uint64()
    exit_cur = C.c_int64()
    shard_start = rank * n_bytes
    gather_i64 = rdist.torch_all_gather_i64(device=dev) if world > 1 else None
    exchange_rounds = [0]
    phase = (C.c_float * 4)()
    step_phase = [0.0, 0.0, 0.0]   # scan, chain, emit of the current step (sharded path: summed over its calls)

    def step():
        if world == 1:
            r = L.rgx_find_all_dev(ctx, pat._h, buf.data_ptr(), n_bytes, -1, d_out.data_ptr(), d_reps.data_ptr(), cap_rec, C.byref(n_rec))
            _lib.check(r)
            return r
        # sharded: scan once, then replay the cursor until every rank's entry == its predecessor's exit
        calls = [0]

        def shard_call(entry_global, mode):
            r = L.rgx_find_all_shard_dev(ctx, pat._h, buf.data_ptr(), n_bytes + halo, n_bytes, int(rank == world - 1),
                                         entry_global - shard_start, shard_start, mode, d_out.data_ptr(),
                                         d_reps.data_ptr(), cap_rec, C.byref(n_rec), C.byref(exit_cur))
            _lib.check(r)
            return r
        last_entry = [0]

        def resolve(entry_global):
            # mode bit 1: cursor replay only; bit 0: the scan of this step is already cached
            shard_call(entry_global, 2 | int(calls[0] > 0))
            L.rgx_ctx_last_timing(ctx, phase)
            if calls[0] == 0:
                step_phase[0], step_phase[1] = phase[0], phase[1]
            else:
                step_phase[1] += phase[1]      # a corrected replay
            calls[0] += 1
            last_entry[0] = entry_global
            return shard_start + exit_cur.value, None

        def finish():
            r = shard_call(last_entry[0], 4)      # output only
            L.rgx_ctx_last_timing(ctx, phase)
            step_phase[2] = phase[2]
            return r
        _, _, total_local, rounds = rdist.resolve_cursor_chain(resolve, rank, world, shard_start, gather_i64, finish=finish,
                                                                   all_starts=[r * n_bytes for r in range(world)])
        exchange_rounds[0] = rounds
        return total_local

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        total_matches = step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = rg.launches(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scan_ms = []
    ev0.record(stream)
    for _ in range(args.steps):
        step()
        if world == 1:
            L.rgx_ctx_last_timing(ctx, phase)
            scan_ms.append((phase[0], phase[1], phase[2]))
        else:
            scan_ms.append(tuple(step_phase))
    ev1.record(stream)
    barrier()
    sampler.stop_flag = True
    ms_total = ev0.elapsed_time(ev1)
    launches = rg.launches(local_rank) - launches0
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        tm = torch.tensor([int(total_matches)], dtype=torch.int64, device=dev)
        dist.all_reduce(tm)
        total_matches = int(tm.item())
    ms_per_step = ms_total / args.steps
    value = world * n_bytes / (ms_per_step * 1e-3) / 1e9

    # end to end through the host-buffer C ABI (rank-local; max over ranks)
    e2e = None
    if not args.no_e2e:
        h_in = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
        h_in.copy_(buf[:n_bytes])
        # result capacity: what the device-resident steps produced plus slack (a rank-local buffer holds about as
        # many records as its shard did); keeps the pinned allocation near 1 GB per rank instead of 5
        cap_e2e = min(cap_rec, int(n_rec.value) + int(n_rec.value) // 8 + 65536)
        h_out = torch.empty(cap_e2e * nc, dtype=torch.int64, pin_memory=True)
        h_reps = torch.empty(cap_e2e, dtype=torch.int32, pin_memory=True)
        torch.cuda.synchronize()

        def e2e_step():
            return _lib.check(L.rgx_find_all_rle(ctx, pat._h, h_in.data_ptr(), n_bytes, -1, h_out.data_ptr(), h_reps.data_ptr(), cap_e2e,
                                                 C.byref(n_rec)))
        k_e2e = max(3, min(args.steps, 5))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            tot = e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / k_e2e
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        if world == 1:
            assert tot == total_matches
        e2e = {"value": world * n_bytes / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": n_bytes,
               "d2h_bytes_per_step": int(n_rec.value) * (nc * 8 + 4) + 256, "ms_per_step": dt * 1e3,
               "api": "rgx_find_all_rle (host buffers, run-length result records)"}
        del h_in, h_out, h_reps

    if rank == 0:
        peak, peak_src = measured_peak()
        scan = float(np.mean([s[0] for s in scan_ms]))
        achieved = n_bytes / (scan * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                traffic = json.load(fh).get(args.workload)
        except Exception:
            pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "kernel": "findall_scan5_kernel<4>" if args.workload == "c3" else "findall_scan_btrun_kernel", "peak_source": peak_src, "scan_ms": scan,
                    "chain_ms": float(np.mean([s[1] for s in scan_ms])), "emit_ms": float(np.mean([s[2] for s in scan_ms])),
                    "algorithmic_bytes_per_launch": n_bytes}
        cpu = None
        if not args.no_cpu:
            cores = os.cpu_count() or 1
            v, dt, passes = CpuOracleSample(wl["kind"], pat.blob(), args.ref_mib, cores).throughput(10.0)
            cpu = {"value": v, "unit": "GB/s", "cores": cores, "kind": "port",
                   "sample": f"{cores} threads x {args.ref_mib} MiB of the same workload x {passes} passes, FindAllBytes(-1), C oracle (not Go)",
                   "seconds": dt}
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": wl["desc"], "bytes_per_gpu": n_bytes, "l2": "input larger than L2, no flush needed",
                       "matches_per_step": int(total_matches), "distinct_records_per_step": int(n_rec.value),
                       "result_form": "run-length offset records left in HBM", "gen_seconds": gen_s,
                       "sharding": None if world == 1 else f"one logical buffer of {world}x{n_bytes} B, 1 MiB halo, exit-cursor all_gather (NCCL), "
                                   f"{exchange_rounds[0]} exchange round(s); e2e is per-rank host buffers"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": sampler.result(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
