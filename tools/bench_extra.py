#!/usr/bin/env python3
"""Side measurements of the two hot-path entry points bench.py does not time (BASELINE.json configs[3], [4]
at single-GPU scale): batched MatchBytes over the corpus patterns and FindReader over a device-resident
stream.  One JSON line per measurement, CUDA events on the library's stream, inputs resident in HBM.

    python tools/bench_extra.py [--inputs-per-pattern 1000000] [--stream-gib 1]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--inputs-per-pattern", type=int, default=1_000_000)
    ap.add_argument("--max-patterns", type=int, default=1000)
    ap.add_argument("--stream-gib", type=float, default=1.0)
    args = ap.parse_args()
    import torch
    import regengo_b200 as rg
    from regengo_b200 import _lib, synth
    L = _lib.load()
    dev = torch.device("cuda", 0)
    ctx = rg.context(0)
    stream = torch.cuda.ExternalStream(L.rgx_ctx_stream(ctx), device=dev)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

    # ---- C4: batched MatchBytes, every ASCII-semantics corpus pattern, seeded mutations of its corpus inputs
    corpus = json.load(open(os.path.join(ROOT, "tests", "golden", "corpus_expected.json")))
    seen, tot_bytes, tot_inputs, tot_ms, n_pat, launches0 = set(), 0, 0, 0.0, 0, rg.launches(0)
    slowest = []
    for ent in corpus["e2e"] + corpus["curated"]:
        pat = ent["pattern"]
        if pat in seen or len(seen) >= args.max_patterns:
            continue
        if any(ord(ch) > 127 for ch in pat) or "\\p{" in pat:
            continue
        seen.add(pat)
        try:
            p = rg.Pattern(pat)
        except Exception:
            continue
        base = synth.mutate_inputs([c["input"] for c in ent["cases"]], 4096, stream=len(seen))
        data, offs = rg.pack_inputs(base)
        reps = max(1, args.inputs_per_pattern // len(base))
        d_data = torch.from_numpy(data).to(dev).repeat(reps)
        lens = np.diff(offs.astype(np.int64))
        all_offs = np.concatenate([[0], np.cumsum(np.tile(lens, reps))]).astype(np.uint64)
        d_offs = torch.from_numpy(all_offs.view(np.int64)).to(dev)
        n = all_offs.size - 1
        d_out = torch.empty(n, dtype=torch.uint8, device=dev)
        for _ in range(2):
            _lib.check(L.rgx_match_batch_dev(ctx, p._h, d_data.data_ptr(), d_offs.data_ptr(), n, d_out.data_ptr()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            _lib.check(L.rgx_match_batch_dev(ctx, p._h, d_data.data_ptr(), d_offs.data_ptr(), n, d_out.data_ptr()))
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        # spot parity of the tiled batch against the first copy
        first = d_out[: len(base)].cpu().numpy()
        assert np.array_equal(first, d_out[len(base): 2 * len(base)].cpu().numpy()) if reps > 1 else True
        tot_bytes += int(all_offs[-1]); tot_inputs += n; tot_ms += ms; n_pat += 1
        slowest.append((ms, pat, p.info.match_engine, int(all_offs[-1])))
        del d_data, d_offs, d_out
    slowest.sort(reverse=True)
    alg = tot_bytes + 8 * tot_inputs + tot_inputs
    if n_pat:
      print(json.dumps({"measure": "C4 batched MatchBytes (rgx_match_batch_dev), corpus patterns, one launch per pattern",
                      "patterns": n_pat, "inputs": tot_inputs, "input_bytes": tot_bytes, "ms": tot_ms,
                      "inputs_per_s": tot_inputs / (tot_ms * 1e-3), "input_GBps": tot_bytes / (tot_ms * 1e-3) / 1e9,
                      "algorithmic_GBps": alg / (tot_ms * 1e-3) / 1e9, "frac_of_measured_hbm": alg / (tot_ms * 1e-3) / 1e9 / peak,
                      "gpu_launches": rg.launches(0) - launches0,
                      "slowest": [{"ms": round(m, 3), "pattern": q, "engine": e, "bytes": b} for m, q, e, b in slowest[:5]]}), flush=True)

    # ---- C5: FindReader over a device-resident stream, default 64 KiB buffer and 1 MiB buffer
    n_bytes = int(args.stream_gib * (1 << 30))
    for name, pattern, kind, kw, sizes in (("DatePattern", synth.DATE_CAPTURE_PATTERN, "stream", dict(digit_noise=0.02), (0, 1 << 20)),
                                           ("URLCapture (TDFA)", synth.URL_PATTERN, "url", {}, (0,))):
      p = rg.Pattern(pattern)
      buf = synth.make_buffer(kind, n_bytes, device=dev, **kw)
      cap = n_bytes // 40
      nc = p.num_cap
      d_so = torch.empty(cap, dtype=torch.int64, device=dev)
      d_ci = torch.empty(cap, dtype=torch.int32, device=dev)
      d_rec = torch.empty(cap * nc, dtype=torch.int64, device=dev)
      for bsz in sizes:
          def call():
              return _lib.check(L.rgx_find_reader_dev(ctx, p._h, buf.data_ptr(), 0, n_bytes, n_bytes, bsz, 0, 0, -1,
                                                      d_so.data_ptr(), d_ci.data_ptr(), d_rec.data_ptr(), cap))
          cnt = call()
          e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          e0.record(stream)
          for _ in range(3):
              call()
          e1.record(stream)
          torch.cuda.synchronize()
          ms = e0.elapsed_time(e1) / 3
          print(json.dumps({"measure": "C5 FindReader (rgx_find_reader_dev), " + name + ", device-resident stream",
                            "buffer_size": bsz or 65536, "stream_bytes": n_bytes, "matches": int(cnt), "ms": ms,
                            "GBps": n_bytes / (ms * 1e-3) / 1e9, "frac_of_measured_hbm": n_bytes / (ms * 1e-3) / 1e9 / peak}), flush=True)


if __name__ == "__main__":
    main()
