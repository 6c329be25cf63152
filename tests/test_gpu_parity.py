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
UNSUPPORTED = set()


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


def test_find_all_sharded_api_single_gpu():
    # the multi-GPU path of bench.py, played on one GPU: three shards of one buffer scanned one after the
    # other with rgx_find_all_shard_dev, the exit cursor of shard r fed to shard r+1
    import ctypes as C
    import torch
    from regengo_b200 import _lib
    from regengo_b200 import dist as rd
    p, o = pair(synth.URL_PATTERN)
    L = _lib.load()
    ctx = rg.context(0)
    total = 3 * synth.BLOCK + 4097
    host = synth.make_buffer("url", total)
    dbuf = torch.from_numpy(host).cuda()
    nc = p.num_cap
    cap = total // 16
    d_out = torch.empty(cap * nc, dtype=torch.int64, device="cuda")
    d_reps = torch.empty(cap, dtype=torch.int32, device="cuda")
    recs, reps = [], []
    entry = 0
    for r in range(3):
        sh = rd.shard_buffer(total, 3, r, halo=65536)
        n_rec, exit_cur = C.c_uint64(), C.c_int64()
        tot = L.rgx_find_all_shard_dev(ctx, p._h, dbuf.data_ptr() + sh.start, sh.buf_len, sh.shard_len, int(sh.is_last),
                                       entry - sh.start, sh.start, 0, d_out.data_ptr(), d_reps.data_ptr(), cap,
                                       C.byref(n_rec), C.byref(exit_cur))
        _lib.check(tot)
        # a second call that only replays the cursor over the cached records gives the same answer
        n_rec2, exit2 = C.c_uint64(), C.c_int64()
        tot2 = L.rgx_find_all_shard_dev(ctx, p._h, dbuf.data_ptr() + sh.start, sh.buf_len, sh.shard_len, int(sh.is_last),
                                        entry - sh.start, sh.start, 1, d_out.data_ptr(), d_reps.data_ptr(), cap,
                                        C.byref(n_rec2), C.byref(exit2))
        assert (tot2, n_rec2.value, exit2.value) == (tot, n_rec.value, exit_cur.value)
        # the multi-GPU exchange: replay from a wrong guess without output (mode 2|1), replay from the true
        # entry (2|1), then output only (4) -- same records, same exit
        wrong = C.c_int64()
        _lib.check(L.rgx_find_all_shard_dev(ctx, p._h, dbuf.data_ptr() + sh.start, sh.buf_len, sh.shard_len, int(sh.is_last),
                                            0, sh.start, 3, d_out.data_ptr(), d_reps.data_ptr(), cap, C.byref(n_rec2), C.byref(wrong)))
        _lib.check(L.rgx_find_all_shard_dev(ctx, p._h, dbuf.data_ptr() + sh.start, sh.buf_len, sh.shard_len, int(sh.is_last),
                                            entry - sh.start, sh.start, 3, d_out.data_ptr(), d_reps.data_ptr(), cap, C.byref(n_rec2), C.byref(exit2)))
        assert n_rec2.value == 0 and exit2.value == exit_cur.value
        d_out.zero_(); d_reps.zero_()
        tot3 = L.rgx_find_all_shard_dev(ctx, p._h, dbuf.data_ptr() + sh.start, sh.buf_len, sh.shard_len, int(sh.is_last),
                                        entry - sh.start, sh.start, 4, d_out.data_ptr(), d_reps.data_ptr(), cap, C.byref(n_rec2), C.byref(exit2))
        assert (tot3, n_rec2.value, exit2.value) == (tot, n_rec.value, exit_cur.value)
        torch.cuda.synchronize()
        recs.append(d_out[: n_rec.value * nc].view(-1, nc).cpu().numpy().copy())
        reps.append(d_reps[: n_rec.value].cpu().numpy().copy())
        assert int(reps[-1].sum()) == tot
        entry = sh.start + exit_cur.value
    expanded = np.repeat(np.concatenate(recs), np.concatenate(reps).astype(np.int64), axis=0)
    en, erecs = o.find_all(host)
    assert expanded.shape[0] == en
    assert np.array_equal(expanded, erecs)


def test_find_all_run_anchor_path():
    # backtracking patterns of the shape C+ b rest take the run-anchor scan (kernels_btrun.cuh)
    cases = [
        (r"(\w+)@(\w)", [b"a@bc@d", b"ab@cd@ef@g", b"@a@b", b"x@", b"aaaa@bbbb@cccc@d" * 30]),
        (synth.EMAIL_PATTERN, [b"a@b.c@d.e", b"user@host.com,other@x.yz;", b"a..b@c.d", b"x@y.z@w.v name@host @ a@.b a@b."]),
        (r"(?P<k>[a-z]+)=(?P<v>\d*);", [b"abc=12;de=;f=3", b"k=v;kk=1;", b"=1;a=2;"]),
        (r"([a-c]+)-(x|[a-c]+-y)", [b"abc-x abc-abc-y ab-ab-x", b"a-a-a-a-y"]),
    ]
    for pat, inputs in cases:
        p, o = pair(pat)
        for b in inputs:
            check_find_all(p, o, b)
            check_find_all(p, o, (b + b" ") * 600)        # crosses 8 KiB segments and 64 KiB parts
            check_find_all(p, o, b * 3, n=2)
    # a run longer than a segment, and one longer than the 4095-byte record field (generic scan fallback)
    p, o = pair(synth.EMAIL_PATTERN)
    check_find_all(p, o, b"x" * 3000 + b"@host.com " + b"y" * 9000 + b"@h.io " + b"z" * 100 + b"@q.rs")
    buf = synth.make_buffer("log", 2 * synth.BLOCK)
    check_find_all(p, o, buf[5:])                        # unaligned device pointer


@pytest.mark.parametrize("pattern,lit", [
    (r"(?P<k>ab)(?P<v>[\w.]+)(?::(?P<n>\d+))?", b"ab"),          # 2-byte literal prefix
    (r"(?P<k>key)=(?P<v>[\w.]+)(?:;(?P<n>\d+))?", b"key="),      # prefix longer than 3, first 4 bytes compared
    (r"(?P<k>id:)(?P<v>\w+)(?:/(?P<n>\d+))?", b"id:"),           # 3-byte literal prefix
    (synth.URL_PATTERN, b"http"),
])
def test_fast_scan_prefix_boundaries(pattern, lit):
    """The fast TDFA scan compares the first <= 4 bytes of the literal prefix 64 bytes per lane, 2 KiB per
    block, 32 KiB per segment: put occurrences (and near misses) across every one of those boundaries and
    at both ends of the buffer, at every alignment of the device pointer the API can produce."""
    p, o = pair(pattern, force_tdfa=True)
    assert p.info.find_engine == 2
    tail = {b"ab": b"x.y:77", b"key=": b"v.1;5", b"id:": b"w9/31", b"http": b"://h.io:8/p"}[lit]
    rng = np.random.default_rng(7)
    n = 3 * 32768 + 700
    buf = bytearray(rng.choice(np.frombuffer(b"abkeyid:=htp xyz\n", dtype=np.uint8), size=n).tobytes())
    tok = lit + tail
    for edge in (64, 128, 2048, 4096, 32768, 65536, 98304):
        for d in range(-len(tok) - 1, 3):
            pos = edge + d
            if 0 <= pos and pos + 1 <= n and rng.integers(0, 2):
                buf[pos:pos + len(tok)] = tok[: n - pos]
    buf[:len(tok)] = tok
    buf[n - len(lit):] = lit                       # a prefix cut off by the end of the buffer
    check_find_all(p, o, bytes(buf))
    check_find_all(p, o, bytes(buf[: n - 1]))
    check_find_all(p, o, bytes(buf[5: 40000]))
    check_find_all(p, o, bytes(buf[: 65536]))        # ends exactly on a segment boundary
    check_find_all(p, o, bytes(buf[: 32768]))
    check_find_all(p, o, bytes(buf[: 32768 + 3]))
    check_find_all(p, o, lit)
    check_find_all(p, o, tok)
    check_find_all(p, o, (tok + b" ") * 3000)       # dense: every lane has several hits


def test_find_all_host_buffer_pipelined_upload():
    """rgx_find_all over a host buffer uploads the input in chunks and scans chunk i as a shard of the buffer
    while chunk i+1 is in flight (entry cursor = previous exit, halo = next chunk).  With 64 KiB chunks a
    3.3 MiB buffer takes 53 chunks; URLs straddle many chunk boundaries."""
    from regengo_b200 import _lib
    p, o = pair(synth.URL_PATTERN)
    L = _lib.load()
    ctx = rg.context(0)
    _lib.check(L.rgx_ctx_set_chunk_bytes(ctx, 1 << 16))
    try:
        buf = bytearray(synth.make_buffer("url", 3 * synth.BLOCK + 300001).tobytes())
        u = b"https://ab.example.com:8080/p/q"
        for k in range(1, 40):
            pos = k * 65536 - (k % len(u))
            buf[pos:pos + len(u)] = u
        check_find_all(p, o, bytes(buf))
        check_find_all(p, o, bytes(buf[: 2 * 65536]))
        check_find_all(p, o, bytes(buf[: 2 * 65536 + 1]))
        # the email pattern is not on the sharded path: the call must take the one-shot route and agree
        p2, o2 = pair(synth.EMAIL_PATTERN)
        check_find_all(p2, o2, synth.make_buffer("log", synth.BLOCK + 17))
    finally:
        _lib.check(L.rgx_ctx_set_chunk_bytes(ctx, 256 << 20))


def test_find_all_at_scale_properties():
    """Larger buffers than the oracle can check everywhere: (1) 32 MiB of each bench workload against the
    oracle, bit for bit; (2) 256 MiB of the URL workload three ways -- one device call, three shards chained
    through their exit cursors, and the pipelined host-buffer call with 16 MiB chunks -- must give identical
    records and repeat counts (the cursor replay crosses every shard / chunk boundary)."""
    import ctypes as C
    import torch
    from regengo_b200 import _lib
    from regengo_b200 import dist as rd
    for kind, pat in (("url", synth.URL_PATTERN), ("log", synth.EMAIL_PATTERN)):
        p, o = pair(pat)
        buf = synth.make_buffer(kind, 32 * synth.BLOCK + 4321, first_block=1000)
        assert check_find_all(p, o, buf) > 50000
    p, _ = pair(synth.URL_PATTERN)
    L = _lib.load()
    ctx = rg.context(0)
    total = 256 * synth.BLOCK + 12345
    dbuf = synth.make_buffer("url", total, first_block=5000, device=torch.device("cuda", 0))
    nc = p.num_cap
    cap = total // 64
    d_out = torch.empty(cap * nc, dtype=torch.int64, device="cuda")
    d_reps = torch.empty(cap, dtype=torch.int32, device="cuda")
    n_rec = C.c_uint64()
    tot = _lib.check(L.rgx_find_all_dev(ctx, p._h, dbuf.data_ptr(), total, -1, d_out.data_ptr(), d_reps.data_ptr(), cap, C.byref(n_rec)))
    torch.cuda.synchronize()
    ref_recs = d_out[: n_rec.value * nc].clone()
    ref_reps = d_reps[: n_rec.value].clone()
    assert int(ref_reps.sum().item()) == tot and n_rec.value > 500000
    # three shards, exit cursor of one is the entry of the next
    entry, got_recs, got_reps, got_tot = 0, [], [], 0
    for r in range(3):
        sh = rd.shard_buffer(total, 3, r, halo=1 << 20)
        n2, exit_cur = C.c_uint64(), C.c_int64()
        t2 = _lib.check(L.rgx_find_all_shard_dev(ctx, p._h, dbuf.data_ptr() + sh.start, sh.buf_len, sh.shard_len, int(sh.is_last),
                                                 entry - sh.start, sh.start, 0, d_out.data_ptr(), d_reps.data_ptr(), cap,
                                                 C.byref(n2), C.byref(exit_cur)))
        torch.cuda.synchronize()
        got_recs.append(d_out[: n2.value * nc].clone()); got_reps.append(d_reps[: n2.value].clone()); got_tot += t2
        entry = sh.start + exit_cur.value
    assert got_tot == tot
    assert torch.equal(torch.cat(got_recs), ref_recs) and torch.equal(torch.cat(got_reps), ref_reps)
    # pipelined host-buffer call
    host = dbuf.cpu().numpy()
    h_out = np.empty(cap * nc, dtype=np.int64)
    h_reps = np.empty(cap, dtype=np.uint32)
    _lib.check(L.rgx_ctx_set_chunk_bytes(ctx, 16 << 20))
    try:
        t3 = _lib.check(L.rgx_find_all_rle(ctx, p._h, host.ctypes.data, total, -1, h_out.ctypes.data, h_reps.ctypes.data, cap, C.byref(n_rec)))
    finally:
        _lib.check(L.rgx_ctx_set_chunk_bytes(ctx, 256 << 20))
    assert t3 == tot and n_rec.value == ref_reps.numel()
    assert np.array_equal(h_out[: n_rec.value * nc], ref_recs.cpu().numpy())
    assert np.array_equal(h_reps[: n_rec.value].astype(np.int64), ref_reps.cpu().numpy().astype(np.int64))


def test_match_multi_one_launch_for_the_suite():
    """rgx_match_multi: every corpus pattern's inputs in ONE packed batch and one launch; flags equal the oracle's
    MatchBytes per pattern.  Tiles of 256 inputs are cut at program boundaries; some programs get no input at all,
    one gets inputs longer than the tile staging buffer (global-memory fallback), some inputs are empty."""
    pats, oracles, inputs, first = [], [], [], [0]
    k = 0
    for ent in CORPUS["e2e"] + CORPUS["curated"]:
        if ent["pattern"] in UNSUPPORTED:
            continue
        p, o = pair(ent["pattern"])
        pats.append(p); oracles.append(o)
        mine = [c["input"].encode("utf-8") for c in ent["cases"]]
        if k % 7 != 3:                                  # every seventh program has no inputs
            mine += synth.mutate_inputs([c["input"] for c in ent["cases"]], 150 + 37 * (k % 11), stream=1000 + k)
            mine += [b"", b"a"]
        else:
            mine = []
        if k == 5:
            mine += [b"x" * 30000 + mine[0] + b"y" * 7, b"z" * 100]   # one tile's bytes exceed the staging buffer
        inputs += mine
        first.append(len(inputs))
        k += 1
    data, offs = rg.pack_inputs(inputs)
    got = rg.match_multi(pats, data, offs, first)
    assert got.shape[0] == len(inputs) and len(pats) >= 230
    for j, o in enumerate(oracles):
        lo, hi = first[j], first[j + 1]
        if hi > lo:
            exp = o.match_batch(data, offs[lo:hi + 1])
            assert np.array_equal(got[lo:hi], exp), (pats[j].pattern, [inputs[lo + i] for i in np.nonzero(got[lo:hi] != exp)[0][:3]])
    # a sub-range of the program table (prog_first[0] > 0) gives the same flags
    sub = rg.match_multi(pats[10:20], data, offs, first[10:21])
    assert np.array_equal(sub, got[first[10]:first[20]])


def test_find_all_kernel_edge_domains():
    """Edge domains of the FindAll kernels, on the kernels themselves: cursor gaps of 2^22 bytes and more (the chain's
    64-bit path and its float/integer quotient repair), filter hits that touch (the SWAR test's redo path), buffers of
    exactly one, two and three chain parts (worklists of length 0 / 1), a walk longer than the event log."""
    p, o = pair(synth.URL_PATTERN)
    u = b"http://ab.cd/e"
    # gaps >= 2^22 before a match, then short and long match lengths right behind each other
    buf = bytearray(b"x" * ((1 << 22) + 4099)) + u + b" " + bytearray(b"y" * ((1 << 22) + 17)) + b"https://a.bc " + u * 3
    check_find_all(p, o, bytes(buf))
    check_find_all(p, o, b"q" * (9 << 20) + u)                     # one match behind a 9 MiB gap: 674 k repeats of it
    # touching / overlapping prefix occurrences
    for s in (b"hthttp://a.b", b"httphttp://a.b", b"http:/http://a.b/http://c.d", b"hhhhttp://x.yz httpp://q.r http//:", b"http://" * 300):
        check_find_all(p, o, s)
        check_find_all(p, o, s * 700)
    # part geometry: 128 KiB chain parts
    base = synth.make_buffer("url", 3 * (128 << 10) + 64)
    for n in ((128 << 10), (128 << 10) + 1, (256 << 10) - 1, (256 << 10), (256 << 10) + 30, 3 * (128 << 10)):
        check_find_all(p, o, base[:n])
    # a URL with more state changes than the event log holds: host labels and path segments alternate a hundred times
    long_url = b"http://" + b".".join([b"a"] * 60) + b":8080/" + b"/".join([b"b"] * 60)
    check_find_all(p, o, b"see " + long_url + b" and " + u)


def test_fast_scan_first_byte_set_filter():
    """TDFA patterns without a literal first byte whose first-byte SET is one or two ASCII ranges (\\d..., [a-z]...) take
    the fast scan with a SWAR range filter: every byte of the set is a candidate start.  Sparse and dense inputs, set
    bytes at block / segment boundaries and at both ends, bytes >= 128 (never in an ASCII range)."""
    semver = r"(?P<major>\d+)\.(?P<minor>\d+)\.(?P<patch>\d+)(?:-(?P<prerelease>[\w.-]+))?(?:\+(?P<build>[\w.-]+))?"
    ipv4 = r"(?P<ip>(?P<a>\d{1,3})\.(?P<b>\d{1,3})\.(?P<c>\d{1,3})\.(?P<d>\d{1,3}))(?::(?P<port>\d{1,5}))?"
    two = r"(?P<k>[a-cx-z][0-9]+)=(?P<v>\d+)"
    rng = np.random.default_rng(3)
    for pat, toks in ((semver, [b"1.2.3", b"10.20.30-rc.1+build.5", b"7.8", b"2024.1.15"]), (ipv4, [b"10.0.0.1", b"192.168.1.254:8080", b"1.2.3", b"999.1.1.1"]),
                      (two, [b"a1=2", b"z99=100", b"d1=2", b"x5=", b"c0=7"])):
        p, o = pair(pat, force_tdfa=True)
        assert p.info.find_engine == 2 and p.device_plan()["fast_tdfa_scan"] == 1
        words = [b"lorem", b"ipsum", b"v", b"caf\xc3\xa9", b"-", b"release", b"host", b"=", b"x", b"\n"]
        parts = []
        for _ in range(40000):
            parts.append(toks[int(rng.integers(0, len(toks)))] if rng.integers(0, 12) == 0 else words[int(rng.integers(0, len(words)))])
        buf = b" ".join(parts)
        check_find_all(p, o, buf)
        check_find_all(p, o, buf[7:70001])
        for t in toks:
            check_find_all(p, o, t)
            check_find_all(p, o, (t + b" ") * 2500)            # dense: grows slabs or falls back to the generic scan
            arr = bytearray(b"." * (2 * 32768 + 64))
            for edge in (2048, 32768, 65536):
                for d in range(-len(t), 2):
                    if rng.integers(0, 2):
                        arr[edge + d: edge + d + len(t)] = t
            arr[:len(t)] = t
            arr[len(arr) - len(t):] = t
            check_find_all(p, o, bytes(arr))


def test_find_all_straight_line_scan():
    """Straight-line backtracking programs (no Alt, one ASCII class per step) take findall_scan_linear_kernel: the records
    of a segment are the bits of a bit-parallel plane.  Lengths 1, 10 and 32; matches across unit (64 B), iteration
    (2 KiB) and segment (8 KiB) boundaries, at both buffer ends, overlapping candidates (the cursor keeps the leftmost),
    all-digit input (every start is a record: the slabs grow), bytes >= 128, odd buffer lengths, device pointers at
    every 16-byte misalignment."""
    import torch
    import ctypes as C
    from regengo_b200 import _lib
    rng = np.random.default_rng(29)
    date = r"(?P<year>\d{4})-(?P<month>\d{2})-(?P<day>\d{2})"
    for pat in (date, r"(\d)", r"(\d{32})", r"(?P<a>[a-f])(?P<b>\d\d)(x)"):
        p, o = pair(pat)
        assert p.device_plan()["linear_findall_scan"] == 1, pat
        S = p.device_plan()["straight_line_steps"]
        sample = {10: b"2024-01-15", 1: b"7", 32: b"12345678901234567890123456789012", 4: b"c42x"}[S]
        buf = bytearray(rng.choice(np.frombuffer(b"abcxyz -\n\xc3\xa9", dtype=np.uint8), size=5 * 8192 + 333).tobytes())
        for edge in (64, 2048, 8192, 16384, 3 * 8192):
            for d in range(-S, 2, max(1, S // 3)):
                buf[edge + d: edge + d + S] = sample
        buf[:S] = sample
        buf[len(buf) - S:] = sample
        buf[700:700 + 3 * S] = sample * 3                  # adjacent
        buf[900:900 + S - 1] = sample[:S - 1]              # cut short
        check_find_all(p, o, bytes(buf))
        check_find_all(p, o, bytes(buf[:8192 + S - 1]))    # the last match ends one byte past the buffer
        check_find_all(p, o, bytes(buf), 7)                # n limit
        check_find_all(p, o, sample)
        check_find_all(p, o, sample[:-1])
        check_find_all(p, o, b"5" * 40000 if S != 4 else b"c42x" * 10000)
        # device pointers at every misalignment of the 16-byte loads
        data = np.frombuffer(bytes(buf), dtype=np.uint8)
        d_all = torch.from_numpy(np.concatenate([np.zeros(32, np.uint8), data])).cuda()
        nc = p.num_cap
        cap = data.size + 16
        d_out = torch.empty(cap * nc, dtype=torch.int64, device="cuda")
        d_reps = torch.empty(cap, dtype=torch.int32, device="cuda")
        n_rec = C.c_uint64()
        for mis in (1, 5, 8, 15):
            sub = data[mis:]
            ecnt, erecs = o.find_all(sub)
            r = _lib.load().rgx_find_all_dev(rg.context(0), p._h, d_all.data_ptr() + 32 + mis, sub.size, -1, d_out.data_ptr(), d_reps.data_ptr(),
                                             cap, C.byref(n_rec))
            _lib.check(r)
            assert r == ecnt and n_rec.value == ecnt
            assert np.array_equal(d_out[: ecnt * nc].cpu().numpy().reshape(ecnt, nc), erecs), (pat, mis)
            assert bool((d_reps[:ecnt] == 1).all())


def test_find_all_straight_line_prefix_filter():
    """Backtracking programs that START with a straight line of >= 2 single-byte steps (then an Alt, a loop, an
    EmptyWidth ...): the bit plane of those steps is the candidate filter of findall_scan_linear_kernel<false> and the
    goto-machine confirms its set bits.  Candidates that pass the prefix but fail later, matches longer than a unit and
    a segment, prefixes cut by the buffer end, dense candidates."""
    log = (r"(?P<timestamp>\d{4}-\d{2}-\d{2}T\d{2}:\d{2}:\d{2})(?:\.(?P<ms>\d{3}))?(?P<tz>Z|[+-]\d{2}:\d{2})?\s+\[(?P<level>\w+)\]\s+(?P<message>.+)")
    cases = [
        (log, [b"2024-01-15T10:30:00Z [INFO] started ok", b"2024-01-15T10:30:00.123+02:00 [ERROR] failed: x", b"2024-01-15T10:30:00 nope",
               b"2024-01-15T10:30:0", b"1999-12-31T23:59:59   [W] " + b"m" * 9000]),
        (r"(?P<g1>(?P<g2>b{2})+)(?P<g3>c+)", [b"bbc", b"bbbbbbcc", b"bbbc", b"bb", b"bbbbx", b"bb" * 5000 + b"c"]),
        (r"(?P<k>ab)(?P<v>\d+|x)", [b"ab1", b"abx", b"ab", b"aab12", b"abab7"]),
    ]
    rng = np.random.default_rng(31)
    for pat, toks in cases:
        p, o = pair(pat)
        d = p.device_plan()
        assert d["linear_prefix_findall_scan"] == 1 and d["straight_line_prefix_steps"] >= 2, pat
        filler = [b"lorem", b"2024-01", b"ab", b"bb", b"b", b"-", b"T", b"\n", b" ", b"caf\xc3\xa9", b"12:30"]
        parts = []
        for _ in range(30000):
            parts.append(toks[int(rng.integers(0, len(toks)))] if rng.integers(0, 9) == 0 else filler[int(rng.integers(0, len(filler)))])
            parts.append(b"\n" if rng.integers(0, 4) == 0 else b" ")
        buf = b"".join(parts)
        assert check_find_all(p, o, buf) > 100
        check_find_all(p, o, buf[3:200003])
        check_find_all(p, o, buf, 11)
        for t in toks:
            check_find_all(p, o, t)
            arr = bytearray(b"." * (3 * 8192 + 50))
            for edge in (64, 2048, 8192, 16384):
                for dlt in (-len(t), -len(t) // 2, -1, 0):
                    if edge + dlt >= 0 and len(t) < 4000:
                        arr[edge + dlt: edge + dlt + len(t)] = t
            arr[len(arr) - min(len(t), 40):] = t[:min(len(t), 40)]
            check_find_all(p, o, bytes(arr))
