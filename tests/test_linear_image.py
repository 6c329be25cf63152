"""The straight-line packing (csrc/device_program.cu: sl_*) that the bit-plane kernels read -- findall_scan_linear_kernel,
the table-free FindReader chase, the static capture offsets of their records -- executed on the CPU over the image the
library actually packs and compared with the oracle.  No GPU.

A whole straight-line program matches at s exactly when the S bytes at s pass its S class tests, with length S and
every capture at a fixed offset; FindAllBytes is then the leftmost non-overlapping selection over those starts
(find.go:130-316).  A straight-line PREFIX is only a filter: every start the reference matches at must pass it."""
import json
import os

import numpy as np

import regengo_b200 as rg
from oracle import Oracle

from helpers import ROOT

CORPUS = json.load(open(os.path.join(ROOT, "tests", "golden", "corpus_expected.json")))


class Line:
    def __init__(self, pattern):
        self.p = rg.Pattern(pattern)
        self.plan = d = self.p.device_plan()
        words = self.p.device_image()
        self.cm = np.frombuffer(words[d["sl_cm_off"]: d["sl_cm_off"] + 64].tobytes(), dtype=np.uint8)   # cm[c]: bit k = byte c in class k
        self.cls = list(bytes.fromhex(d["sl_cls"]))
        self.cap = list(bytes.fromhex(d["sl_cap"]))
        self.oracle = Oracle(self.p.blob())

    def alive(self, data):
        """alive[s]: the bytes at s pass every step (bytes past the end pass nothing)."""
        a = np.frombuffer(data, dtype=np.uint8)
        n = a.size
        ok = np.ones(n, dtype=bool)
        for i, k in enumerate(self.cls):
            step = np.zeros(n, dtype=bool)
            if i < n:
                step[: n - i] = (self.cm[a[i:]] >> k) & 1
            ok &= step
        return ok


def check_whole(pattern, inputs):
    ln = Line(pattern)
    assert ln.plan["linear_findall_scan"] == 1 and ln.plan["sl_caps_ok"] == 1
    S = ln.plan["straight_line_steps"]
    assert len(ln.cls) == S and len(ln.cap) == ln.p.num_cap and ln.cap[0] == 0 and ln.cap[1] == S
    for data in inputs:
        ok = ln.alive(data)
        recs, cursor = [], 0
        for s in np.nonzero(ok)[0]:
            if s >= cursor:
                recs.append([int(s) + c for c in ln.cap])
                cursor = int(s) + S
        cnt, erecs = ln.oracle.find_all(data)
        assert cnt == len(recs), (pattern, data[:60], cnt, len(recs))
        assert np.array_equal(np.array(recs, dtype=np.int64).reshape(-1, ln.p.num_cap), erecs), (pattern, data[:60])
    return len(inputs)


def test_whole_straight_line_programs_on_the_cpu():
    rng = np.random.default_rng(41)
    alphabet = np.frombuffer(b"0123456789--ab cfx\n\xc3", dtype=np.uint8)
    rand = [rng.choice(alphabet, size=int(n)).tobytes() for n in rng.integers(0, 400, size=60)]
    fixed = [b"", b"2024-01-15", b"2024-01-1", b"x2024-01-15y2024-01-152024-01-15", b"12024-01-15", b"7", b"c42x" * 3, b"1" * 70]
    for pat in (r"(?P<year>\d{4})-(?P<month>\d{2})-(?P<day>\d{2})", r"(\d{4}-\d{2}-\d{2})", r"(\d)", r"(\d{32})", r"(?P<a>[a-f])(?P<b>\d\d)(x)",
                r"(a)(b)(c)", r"([ -~])(\d)"):
        check_whole(pat, fixed + rand)
    # a class that reaches beyond ASCII is not a straight-line step (bytes >= 128 go through the UTF-8 decoder)
    assert rg.Pattern(r"([^\n])(\d)").device_plan()["linear_findall_scan"] == 0
    # every corpus pattern that takes the straight-line scan, on its own inputs
    n = 0
    for ent in CORPUS["e2e"] + CORPUS["curated"]:
        if ent["n_groups"] == 0:
            continue
        try:
            plan = rg.Pattern(ent["pattern"]).device_plan()
        except rg.RegengoError:
            continue
        if plan["linear_findall_scan"]:
            n += check_whole(ent["pattern"], [c["input"].encode("utf-8") for c in ent["cases"]])
    assert n >= 10


def test_straight_line_prefix_is_a_necessary_condition():
    n_pat = 0
    for ent in CORPUS["e2e"] + CORPUS["curated"]:
        if ent["n_groups"] == 0:
            continue
        try:
            ln = Line(ent["pattern"])
        except rg.RegengoError:
            continue
        if not ln.plan["linear_prefix_findall_scan"]:
            continue
        assert 2 <= ln.plan["straight_line_prefix_steps"] == len(ln.cls) <= 32
        for c in ent["cases"]:
            data = c["input"].encode("utf-8")
            ok = ln.alive(data)
            cnt, recs = ln.oracle.find_all(data)
            for r in recs:
                assert ok[int(r[0])], (ent["pattern"], data, int(r[0]))
            # ... and FindBytes' match (found with the skip-restart rule) starts at a position that passes it as well
            f = ln.oracle.find(data)
            if f is not None:
                assert ok[int(f[0])], (ent["pattern"], data)
        n_pat += 1
    assert n_pat >= 5
