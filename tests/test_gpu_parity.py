"""Parity tests proper: the CUDA path, called through the C ABI (regengo_b200.Pattern ->
libregengo_b200.so), against the CPU oracle on the same inputs.  Bit-exact: match booleans and every
capture offset.  Needs a GPU."""
import json
import os

import numpy as np
import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu

import regengo_b200 as rg  # noqa: E402
from regengo_b200 import synth  # noqa: E402
from oracle import Oracle  # noqa: E402

with open(os.path.join(ROOT, "tests", "golden", "corpus_expected.json")) as _fh:
    CORPUS = json.load(_fh)
UNSUPPORTED = {r"\p{L}+", r"\p{Greek}+", r"[\p{L}\p{N}]+", r"\p{Hebrew}+"}


def pair(pattern, **kw):
    p = rg.Pattern(pattern, **kw)
    return p, Oracle(p.blob())


def check_batch(p, o, inputs):
    data, offs = rg.pack_inputs(inputs)
    got = p.match_batch(data, offs)
    exp = o.match_batch(data, offs)
    assert np.array_equal(got, exp), [(inputs[i], int(got[i]), int(exp[i])) for i in np.nonzero(got != exp)[0][:5]]
    if p.info.find_engine:
        gf, gr = p.find_batch(data, offs)
        ef, er = o.find_batch(data, offs)
        assert np.array_equal(gf, ef), [(inputs[i], int(gf[i]), int(ef[i])) for i in np.nonzero(gf != ef)[0][:5]]
        m = ef.astype(bool)
        assert np.array_equal(gr[m], er[m]), [(inputs[i], gr[i].tolist(), er[i].tolist()) for i in np.nonzero((gr != er).any(1) & m)[0][:5]]


def check_find_all(p, o, buf, n=-1):
    cnt, recs = p.find_all_offsets(buf, n)
    ecnt, erecs = o.find_all(buf, n)
    assert cnt == ecnt, (cnt, ecnt)
    assert np.array_equal(recs, erecs), np.nonzero((recs != erecs).any(1))[0][:5]
    return cnt


def test_c1_date_match_10k():
    # BASELINE.json configs[0]: Date MatchBytes over 10k synthetic strings (fires the skip-restart rule)
    p, o = pair(synth.DATE_PATTERN)
    inputs = synth.date_strings(10000)
    data, offs = rg.pack_inputs(inputs)
    got = p.match_batch(data, offs)
    exp = o.match_batch(data, offs)
    assert np.array_equal(got, exp)
    assert 2000 < int(exp.sum()) < 9000
    assert not p.match_bytes(b"12024-01-15") and p.match_bytes(b"99 2024-01-15")
    assert rg.launches() > 0


def test_corpus_batch_match_and_find():
    # every corpus pattern: its own inputs plus seeded mutations, MatchBytes + FindBytes
    n_pat = 0
    for ent in CORPUS["e2e"] + CORPUS["curated"]:
        if ent["pattern"] in UNSUPPORTED:
            continue
        p, o = pair(ent["pattern"])
        inputs = [c["input"].encode("utf-8") for c in ent["cases"]]
        inputs += synth.mutate_inputs([c["input"] for c in ent["cases"]], 200, stream=n_pat)
        inputs += [b"", b"a", b"\n"]
        check_batch(p, o, inputs)
        n_pat += 1
    assert n_pat >= 230


def test_corpus_find_all():
    for ent in CORPUS["e2e"] + CORPUS["curated"]:
        if ent["pattern"] in UNSUPPORTED or ent["n_groups"] == 0:
            continue
        p, o = pair(ent["pattern"])
        for c in ent["cases"]:
            b = c["input"].encode("utf-8")
            check_find_all(p, o, b)
        joined = b" ".join(c["input"].encode("utf-8") for c in ent["cases"]) * 3
        check_find_all(p, o, joined)


def test_forced_engines():
    for kw in ({"force_thompson": True}, {"force_tnfa": True}, {"force_tdfa": True}):
        for pat in (r"(\d+)-(\d+)", r"(?P<k>\w+)=(?P<v>[^;]*);", r"(a|ab)(c|bcd)(d*)"):
            p, o = pair(pat, **kw)
            inputs = [b"12-34", b"x=1;y=22;", b"abcd", b"k=;", b"", b"9-", b"abcdddd abcd"]
            check_batch(p, o, inputs)
            for b in inputs:
                check_find_all(p, o, b * 5)


@pytest.mark.parametrize("kind,pattern,nblocks", [("url", synth.URL_PATTERN, 3), ("log", synth.EMAIL_PATTERN, 2)])
def test_find_all_synthetic_buffer(kind, pattern, nblocks):
    # configs[1]/[2] at oracle-friendly size: same generator, same seed, blocks 0..n
    p, o = pair(pattern)
    buf = synth.make_buffer(kind, nblocks * synth.BLOCK + 12345)
    cnt = check_find_all(p, o, buf)
    assert cnt > 1000
    # limits and odd slices (unaligned device pointers are exercised through the dev API below)
    check_find_all(p, o, buf[: 70000], n=17)
    check_find_all(p, o, buf[3: 8192 + 5])
    check_find_all(p, o, buf[: 1])
    check_find_all(p, o, buf[: 0])


def test_find_all_q2_duplicates_and_segment_straddle():
    p, o = pair(synth.URL_PATTERN)
    check_find_all(p, o, b"z" * 24 + b" http://a.com")
    # URLs straddling the 8 KiB segment and 256 KiB part boundaries, long gaps, adjacent URLs
    buf = bytearray(b"x" * (600 * 1024))
    for pos in (8185, 8192 * 3 - 1, 262140, 262144, 300000, 300020, 524287, 600 * 1024 - 14):
        u = b"https://ab.example.com:8080/p/q"
        buf[pos:pos + len(u)] = u[: len(buf) - pos]
    check_find_all(p, o, bytes(buf))
    check_find_all(p, o, b"http://a.b http://a.b/http://c.d/ http://e.f" * 50)


def test_find_all_dense_matches_grow_slabs():
    p, o = pair(r"(a)")
    check_find_all(p, o, b"a" * 50000 + b"b" * 100 + b"a" * 3000)
    p, o = pair(r"(\w+)")
    check_find_all(p, o, b"ab cd " * 20000)


def test_find_all_nullable_anchored_memo_paths():
    for pat in (r"(a*)", r"^(\w+)", r"(\w+)$", r"(?P<outer>(?P<inner>a+)+)b", r"(x*)$", r"(a|b)*c(d?)"):
        p, o = pair(pat)
        for b in (b"", b"baaab", b"aaaa", b"hello world", b"aaab aab b", b"bbb", b"abcabcd cd"):
            check_find_all(p, o, b)
            check_find_all(p, o, b * 40)


def test_unsupported_and_errors():
    p = rg.Pattern(r"\d+")
    with pytest.raises(rg.RegengoError):
        p.find_all_offsets(b"123")          # no capture groups => no Find* methods (regengo.go:110)
    with pytest.raises(rg.RegengoError):
        rg.Pattern(r"(unclosed")
