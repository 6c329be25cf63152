#!/usr/bin/env python3
"""Small drivers for ncu captures of the kernels bench.py's four workloads do not launch:
    python tools/prof_misc.py linear|generic|table|replace [MiB]
linear : findall_scan_linear_kernel<true>  (DateCapture FindAllBytes)
generic: findall_scan_kernel<FIND_BT>      ((cat|dog)fish FindAllBytes)
table  : find_reader_table_kernel + table-driven chase (a TDFA pattern through FindReader)
replace: replace_batch_kernel<0/1>         (email redaction over log lines)"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import regengo_b200 as rg  # noqa: E402
from regengo_b200 import _lib, synth  # noqa: E402

which = sys.argv[1]
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 128
L = _lib.load()
ctx = rg.context(0)
n = mib << 20
if which in ("linear", "generic"):
    pat = r"(?P<year>\d{4})-(?P<month>\d{2})-(?P<day>\d{2})" if which == "linear" else r"(cat|dog)fish"
    p = rg.Pattern(pat)
    buf = synth.make_buffer("stream" if which == "linear" else "url", n, device=torch.device("cuda", 0))
    nc = p.num_cap
    cap = n // 16 + 1024
    d_out = torch.empty(cap * nc, dtype=torch.int64, device="cuda")
    d_reps = torch.empty(cap, dtype=torch.int32, device="cuda")
    n_rec = C.c_uint64()
    for _ in range(3):
        r = L.rgx_find_all_dev(ctx, p._h, buf.data_ptr(), n, -1, d_out.data_ptr(), d_reps.data_ptr(), cap, C.byref(n_rec))
        _lib.check(r)
    print(which, pat, "matches", r, p.device_plan())
elif which == "table":
    p = rg.Pattern(synth.URL_PATTERN)
    buf = synth.make_buffer("url", n, device=torch.device("cuda", 0))
    nc = p.num_cap
    cap = n // 64 + 1024
    so = torch.empty(cap, dtype=torch.int64, device="cuda")
    ci = torch.empty(cap, dtype=torch.int32, device="cuda")
    rec = torch.empty(cap * nc, dtype=torch.int64, device="cuda")
    for _ in range(3):
        r = L.rgx_find_reader_dev(ctx, p._h, buf.data_ptr(), 0, n, n, 0, 0, 0, -1, so.data_ptr(), ci.data_ptr(), rec.data_ptr(), cap)
        _lib.check(r)
    print("table-driven FindReader, URL pattern, matches", r)
else:
    p = rg.Pattern(synth.EMAIL_PATTERN)
    lines = bytes(synth.make_buffer("log", n)).split(b"\n")
    data, offs = rg.pack_inputs(lines)
    d_data = torch.from_numpy(data.copy()).cuda()
    d_offs = torch.from_numpy(offs.view(np.int64).copy()).cuda()
    d_out = torch.empty(data.size + 64 * len(lines), dtype=torch.uint8, device="cuda")
    d_out_offs = torch.empty(len(lines) + 1, dtype=torch.int64, device="cuda")
    total = C.c_uint64()
    t = b"$user@REDACTED.$tld"
    for _ in range(3):
        _lib.check(L.rgx_replace_batch_dev(ctx, p._h, t, len(t), d_data.data_ptr(), d_offs.data_ptr(), len(lines), d_out.data_ptr(), d_out.numel(),
                                           d_out_offs.data_ptr(), C.byref(total)))
    print("replace", len(lines), "lines", data.size, "->", total.value, "bytes")
