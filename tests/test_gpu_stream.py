"""stream.FindReader parity: the chunk-parallel CUDA path (through the C ABI) against the oracle's
literal emulation of the reference loop (leftover carry, deferral, bytes.Index).  Needs a GPU."""
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import regengo_b200 as rg  # noqa: E402
from regengo_b200 import synth  # noqa: E402
from oracle import Oracle  # noqa: E402


def pair(pattern, **kw):
    p = rg.Pattern(pattern, **kw)
    return p, Oracle(p.blob())


def check_reader(p, o, data, buffer_size=0, max_leftover=0):
    n, so, ci, recs = p.find_reader_offsets(data, rg.StreamConfig(buffer_size, max_leftover))
    en, eso, eci, erecs = o.find_reader(data, buffer_size, max_leftover)
    assert n == en, (n, en)
    assert np.array_equal(so, eso)
    assert np.array_equal(ci, eci)
    assert np.array_equal(recs, erecs)
    return n


def test_reference_stream_kats():
    # tests/integration/streaming/streaming_test.go:190-316 through the device path
    p, o = pair(synth.DATE_CAPTURE_PATTERN)
    pos = [100, 32768, 65530, 65550, 70000, 99000]
    dates = [b"2024-01-01", b"2024-02-02", b"2024-03-03", b"2024-04-04", b"2024-05-05", b"2024-06-06"]
    buf = bytearray(b"x" * (100 * 1024))
    for q, d in zip(pos, dates):
        buf[q:q + 10] = d
    n, so, ci, recs = p.find_reader_offsets(bytes(buf), rg.StreamConfig(64 * 1024, 0))
    assert n == 6 and so.tolist() == pos and ci.tolist() == [0, 0, 1, 1, 1, 1]
    check_reader(p, o, bytes(buf), 64 * 1024)
    data = b"prefix 2024-01-15 middle 2024-02-20 suffix"
    n, so, _, _ = p.find_reader_offsets(data)
    assert so.tolist() == [7, 25]
    # early termination is the callback's business (streaming_test.go:143-162)
    seen = []
    p.find_reader(io.BytesIO(b"2024-01-01 " * 100), rg.StreamConfig(), lambda m: (seen.append(m.stream_offset) or len(seen) < 5))
    assert seen == [0, 11, 22, 33, 44]
    assert p.find_reader_count(io.BytesIO(b"x" * 10000)) == 0
    assert p.find_reader_count(io.BytesIO(b"")) == 0
    with pytest.raises(rg.BufferTooSmall):
        p.find_reader_offsets(b"abc", rg.StreamConfig(1000, 0))


@pytest.mark.parametrize("digit_noise", [0.0, 0.02])
@pytest.mark.parametrize("bufsize,leftover", [(0, 0), (65536, 4096), (1 << 17, 0)])
def test_patterned_stream(digit_noise, bufsize, leftover):
    # configs[4] at oracle-friendly size, with and without the digit noise that fires the skip-restart rule
    p, o = pair(synth.DATE_CAPTURE_PATTERN)
    buf = synth.make_buffer("stream", 3 * synth.BLOCK + 4321, digit_noise=digit_noise)
    n = check_reader(p, o, buf, bufsize, leftover)
    assert n > 50000


def test_exact_fill_flush_pass_and_short_streams():
    p, o = pair(synth.DATE_CAPTURE_PATTERN)
    B, L = 65536, 1024
    stride = B - L
    base = synth.make_buffer("stream", 4 * B, digit_noise=0.02)
    for total in (B, B + stride, B + 2 * stride, B - 1, B + 1, L, L + 1, 10, 1, B + stride - 1, 2 * B):
        check_reader(p, o, base[:total], B, L)


def test_q15_q16_quirks_on_device():
    p, o = pair(synth.DATE_CAPTURE_PATTERN)
    buf = bytearray(b"x" * (100 * 1024))
    buf[64505:64515] = b"2024-01-15"
    assert check_reader(p, o, bytes(buf), 65536) == 0          # straddles the deferral line: dropped
    assert check_reader(p, o, b"12024-01-15 2024-01-15") == 2   # bytes.Index reports the skipped text first


def test_tdfa_and_backtracking_patterns_streaming():
    for pat, kind in ((synth.URL_PATTERN, "url"), (synth.EMAIL_PATTERN, "log"), (r"(\d+)", "stream")):
        p, o = pair(pat)
        buf = synth.make_buffer(kind, 300_000)
        check_reader(p, o, buf)
        check_reader(p, o, buf, 65536, 100)


def test_chunk_range_sharding_matches_whole():
    # the multi-GPU split: chunk ranges processed separately concatenate to the whole result
    p, o = pair(synth.DATE_CAPTURE_PATTERN)
    buf = synth.make_buffer("stream", 2 * synth.BLOCK + 99, digit_noise=0.02)
    n, so, ci, recs = p.find_reader_offsets(buf)
    parts = [p.find_reader_offsets(buf, None, first, cnt) for first, cnt in ((0, 7), (7, 13), (20, -1))]
    assert sum(x[0] for x in parts) == n
    assert np.array_equal(np.concatenate([x[1] for x in parts]), so)
    assert np.array_equal(np.concatenate([x[2] for x in parts]), ci)
    assert np.array_equal(np.concatenate([x[3] for x in parts]), recs)


def test_lane_parallel_chase_boundaries():
    """The chase processes a chunk in tiles of 2 KiB, every lane of a warp starting at the first sync point of its 64
    positions (kernels_stream.cuh, TILE CHASE).  Put the cases that couple lanes and tiles exactly where they meet
    (multiples of 2 KiB, and of 32 KiB for the 1 MiB buffer): dates straddling them, skip-restart prefixes, a skipped copy
    of a date's text far before the real match (bytes.Index reports the copy: an EVENT that restarts the tiles behind
    it), digit runs longer than a tile (no sync point there), and heavy digit noise."""
    p, o = pair(synth.DATE_CAPTURE_PATTERN)
    rng = np.random.default_rng(11)
    for bufsize, step in ((0, 2048), (1 << 20, 32768)):
        n = 3 * (bufsize or 65536) + 777
        buf = bytearray(rng.choice(np.frombuffer(b"abcdefgh \n", dtype=np.uint8), size=n).tobytes())
        for edge in range(step, n - 64, step):
            kind = (edge // step) % 6
            if kind == 0:
                buf[edge - 4:edge + 6] = b"2024-01-15"                  # straddles the nominal range start
            elif kind == 1:
                buf[edge - 3:edge + 8] = b"12024-02-16"                 # skip-restart prefix across it
            elif kind == 2:
                buf[edge - 300:edge + 300] = b"7" * 600                 # digits all over the sync window
            elif kind == 3:
                buf[edge + 250:edge + 260] = b"2031-12-31"              # right behind the cover window
            elif kind == 4:
                buf[edge - 1:edge + 10] = b"-2024-03-17"
        check_reader(p, o, bytes(buf), bufsize)
    # a skipped copy of the text, then the real match several ranges later and nothing in between
    buf = bytearray(b"x" * 70000)
    buf[100:111] = b"12024-01-15"
    buf[9000:9010] = b"2024-01-15"
    buf[30000:30010] = b"2024-01-15"
    assert check_reader(p, o, bytes(buf)) >= 3
    # digit run longer than several ranges, dates inside and after it
    buf = bytearray(b"y" * 200000)
    buf[3000:15000] = b"5" * 12000
    buf[9000:9010] = b"2024-05-05"
    buf[15003:15013] = b"2024-06-06"
    check_reader(p, o, bytes(buf))
    # heavy digit / dash noise
    for noise in (0.2, 0.5):
        check_reader(p, o, synth.make_buffer("stream", 2 * synth.BLOCK + 17, digit_noise=noise))
        check_reader(p, o, synth.make_buffer("stream", 2 * synth.BLOCK + 17, digit_noise=noise), 1 << 20)


def test_tile_chase_events_engines_and_alignment():
    """More of the tile chase: (a) relocation events in bulk -- every date preceded by a digit is jumped over by the
    skip-restart rule and is then a COPY of a later identical date's text, so bytes.Index reports matches there and the
    tiles restart behind them, many times per chunk; (b) straight-line programs of length 1 and 32 (table-free bit-plane
    path) next to table-driven ones (an Alt pattern, a TDFA pattern); (c) leftovers that make the chunk stride odd, so
    chunks start at every alignment of the 16-byte loads; (d) deferral of a match at a tile edge."""
    rng = np.random.default_rng(23)
    p, o = pair(synth.DATE_CAPTURE_PATTERN)
    # (a) few distinct dates, a third of them behind a digit
    dates = [b"2024-01-15", b"1999-12-31"]
    parts = []
    for i in range(30000):
        d = dates[int(rng.integers(0, 2))]
        parts.append((b"7" if rng.integers(0, 3) == 0 else b"") + d + b"x" * int(rng.integers(1, 40)))
    buf = b"".join(parts)
    n = check_reader(p, o, buf)
    assert n > 20000
    check_reader(p, o, buf, 1 << 20)
    # (c) odd strides: BufferSize 64 KiB with leftovers 1001..1016 -> chunk starts at all 16 alignments
    for lo in (1001, 1003, 1008, 1013):
        check_reader(p, o, buf[: 5 * 65536 + 99], 65536, lo)
    # (b) straight-line programs of length 1 and 32; table-driven programs through the same tiles
    text = synth.make_buffer("stream", 3 * synth.BLOCK // 2 + 5, digit_noise=0.05)
    for pat in (r"(\d)", r"(\d{32})", r"(\d{4})-(\d{2})", r"(\d\d-|-\d\d)", r"(?P<y>\d{4})-(?P<m>\d+)"):
        pp, oo = pair(pat)
        check_reader(pp, oo, text)
        check_reader(pp, oo, text, 1 << 20)
    long_digits = bytearray(text[:300000])
    long_digits[5000:5100] = b"1234567890" * 10
    long_digits[70000:70040] = b"9" * 40
    pp, oo = pair(r"(\d{32})")
    assert check_reader(pp, oo, bytes(long_digits)) >= 3
    # (d) a date ending exactly on / one past the deferral line of the first chunk, which is also near a tile edge
    pl, ol = pair(synth.DATE_CAPTURE_PATTERN)
    B, L = ol.stream_config(0, 0)
    for shift in (-1, 0, 1, 2):
        b2 = bytearray(b"z" * (2 * B))
        end = B - L + shift
        b2[end - 10:end] = b"2024-07-04"
        b2[2048 - 5:2048 + 5] = b"2024-08-05"
        check_reader(pl, ol, bytes(b2))


def test_runs_of_chunks_host_pieces_and_device_runs(monkeypatch):
    """Long streams are searched in runs of whole chunks: the host path uploads / searches / downloads them as pieces on
    three streams, the device path bounds the per-chunk hit regions.  RGX_READER_MIN_RUN (a test knob) forces runs of
    1, 3 and 7 chunks on a stream of 40 chunks -- including the flush pass of an exactly filled last buffer -- and the
    records must come out in chunk order, identical to the oracle's."""
    import ctypes as C
    import torch
    from regengo_b200 import _lib
    L = _lib.load()
    p, o = pair(synth.DATE_CAPTURE_PATTERN)
    pu, ou = pair(synth.URL_PATTERN)
    B, Lo = o.stream_config(0, 0)
    date_stream = synth.make_buffer("stream", 40 * (B - Lo) + 1234, digit_noise=0.02)
    exact = bytes(date_stream[: B + 5 * (B - Lo)])                 # the last read fills the buffer exactly: flush pass
    url_stream = synth.make_buffer("url", 12 * 65536 + 999)
    for run in ("1", "3", "7"):
        monkeypatch.setenv("RGX_READER_MIN_RUN", run)
        check_reader(p, o, date_stream)
        check_reader(p, o, exact)
        check_reader(pu, ou, url_stream)
        # device pointers: rgx_find_reader_dev over the whole stream and over a chunk range
        data = np.frombuffer(bytes(date_stream), dtype=np.uint8)
        en, eso, eci, erecs = o.find_reader(data)
        d = torch.from_numpy(data.copy()).cuda()
        nc = p.num_cap
        so = torch.empty(en + 8, dtype=torch.int64, device="cuda")
        ci = torch.empty(en + 8, dtype=torch.int32, device="cuda")
        rec = torch.empty((en + 8) * nc, dtype=torch.int64, device="cuda")
        n = L.rgx_find_reader_dev(rg.context(0), p._h, d.data_ptr(), 0, data.size, data.size, 0, 0, 0, -1, so.data_ptr(), ci.data_ptr(), rec.data_ptr(), en + 8)
        _lib.check(n)
        assert n == en and np.array_equal(so[:en].cpu().numpy(), eso) and np.array_equal(ci[:en].cpu().numpy(), eci)
        assert np.array_equal(rec[: en * nc].cpu().numpy().reshape(en, nc), erecs)
        n2 = L.rgx_find_reader_dev(rg.context(0), p._h, d.data_ptr(), 0, data.size, data.size, 0, 0, 5, 11, so.data_ptr(), ci.data_ptr(), rec.data_ptr(), en + 8)
        _lib.check(n2)
        sel = (eci >= 5) & (eci < 16)
        assert n2 == int(sel.sum()) and np.array_equal(so[:n2].cpu().numpy(), eso[sel]) and np.array_equal(ci[:n2].cpu().numpy(), eci[sel])


def test_find_reader_first_and_count():
    # streaming_test.go:143-162 (early termination), FindReaderCount / FindReaderFirst (streaming.go:258-313)
    import io
    p, o = pair(synth.DATE_CAPTURE_PATTERN)
    data = b"".join(b"entry %d on 2024-%02d-15; " % (i, 1 + i % 12) for i in range(100))
    assert p.find_reader_count(io.BytesIO(data)) == 100
    first, off = p.find_reader_first(io.BytesIO(data))
    assert first.match == b"2024-01-15" and off == data.index(b"2024-01-15")
    assert p.find_reader_first(io.BytesIO(b"x" * 10000)) == (None, 0)
    seen = []
    p.find_reader(io.BytesIO(data), rg.StreamConfig(), lambda m: (seen.append(m.stream_offset), len(seen) < 5)[1])
    assert len(seen) == 5
