#!/usr/bin/env python3
"""Per-pattern time of the batched-MatchBytes suite (c4): which programs dominate the multi-program launch."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import regengo_b200 as rg  # noqa: E402
from regengo_b200 import _lib  # noqa: E402

L = _lib.load()
dev = torch.device("cuda", 0)
ctx = rg.context(0)
stream = torch.cuda.ExternalStream(L.rgx_ctx_stream(ctx), device=dev)
pools = bench.suite_pools()
pats = [rg.Pattern(p) for p, _ in pools]
data, offs, first = bench.suite_batch(pools, int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000, 0, device=dev)
d_offs = torch.from_numpy(offs.astype(np.uint32).view(np.int32)).to(dev)
d_out = torch.empty(int(first[-1]), dtype=torch.uint8, device=dev)
rows = []
for k, p in enumerate(pats):
    h = (C.c_void_p * 1)(p._h)
    pf = np.array([first[k], first[k + 1]], dtype=np.uint64)
    for _ in range(2):
        _lib.check(L.rgx_match_multi_dev(ctx, h, 1, data.data_ptr(), d_offs.data_ptr(), pf.ctypes.data, d_out.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        _lib.check(L.rgx_match_multi_dev(ctx, h, 1, data.data_ptr(), d_offs.data_ptr(), pf.ctypes.data, d_out.data_ptr()))
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    nb = int(offs[int(first[k + 1])] - offs[int(first[k])])
    rows.append((ms, p.pattern, p.info.match_engine, p.info.match_memo, nb, int(first[k + 1] - first[k])))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("total ms", tot, "patterns", len(rows))
for ms, pat, eng, memo, nb, n in rows[:25]:
    print("%8.3f ms  %5.1f%%  engine %d memo %d  %6.1f B/input  %.2f G inputs/s  %s" % (ms, 100 * ms / tot, eng, memo, nb / n, n / ms / 1e6, pat[:70]))
print("median ms", rows[len(rows) // 2][0])
