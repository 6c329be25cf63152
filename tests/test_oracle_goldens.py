"""Oracle pin #1 -- independent of the product's front-end.

tests/golden/golden_blob.py writes program blobs straight from what the reference emitted into its 24 checked-in
generated Go files (instruction listings, class byte sets, Thompson masks, TDFA tables: generated_goldens.json).  The
oracle, fed THOSE programs, must give
  * on the reference's curated inputs (scripts/curated/cases.go, via corpus_expected.json) the results its generated
    tests assert (internal/compiler/test_gen.go:72-239): MatchBytes / FindBytes / FindAllBytes(-1) texts;
  * on the streaming test patterns the absolute stream offsets of tests/integration/streaming/streaming_test.go:190-316.
No line of regengo_b200 runs here; what is pinned is the oracle's executors (oracle/rgx_oracle.c)."""
import json
import os
import sys

import numpy as np
import pytest

from helpers import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import golden_blob  # noqa: E402
from oracle import Oracle  # noqa: E402

GOLDENS = golden_blob.load_goldens()
with open(os.path.join(ROOT, "tests", "golden", "corpus_expected.json")) as _fh:
    CURATED = {e["file"]: e for e in json.load(_fh)["curated"]}


def texts(b, rec):
    return [(b[rec[2 * g]:rec[2 * g + 1]].decode("utf-8", "replace") if rec[2 * g] >= 0 else "") for g in range(len(rec) // 2)]


def test_every_curated_file_has_a_golden():
    files = {e["file"] for e in GOLDENS}
    assert set(CURATED) <= files and len(CURATED) == 17


@pytest.mark.parametrize("idx", range(len(GOLDENS)))
def test_oracle_on_the_reference_s_own_program(idx):
    e = GOLDENS[idx]
    o = Oracle(golden_blob.blob_from_golden(e))
    ent = CURATED.get(e["file"])
    if ent is None:
        pytest.skip("streaming test pattern: covered by the stream KATs below")
    assert e["pattern"] == ent["pattern"]
    for c in ent["cases"]:
        b = c["input"].encode("utf-8")
        assert o.match(b) == c["match"], (e["file"], c["input"])
        if ent["n_groups"] > 0:
            r = o.find(b)
            assert (None if r is None else texts(b, r)) == c["find"], (e["file"], c["input"])
            n, recs = o.find_all(b)
            assert [texts(b, list(x)) for x in recs] == c["findall"], (e["file"], c["input"])


def test_stream_kats_on_the_reference_s_date_program():
    # tests/integration/streaming/streaming_test.go:190-280 (TestStreamingLargeInputBoundary) and :283-316 (TestStreamingOffsets)
    e = next(x for x in GOLDENS if x["file"].endswith("streaming/testdata/date_pattern.go") or x["name"] == "DatePattern")
    o = Oracle(golden_blob.blob_from_golden(e))
    pos = [100, 32768, 65530, 65550, 70000, 99000]
    dates = [b"2024-01-01", b"2024-02-02", b"2024-03-03", b"2024-04-04", b"2024-05-05", b"2024-06-06"]
    buf = bytearray(b"x" * (100 * 1024))
    for q, d in zip(pos, dates):
        buf[q:q + 10] = d
    n, so, ci, recs = o.find_reader(bytes(buf), 64 * 1024, 0)
    assert n == 6 and so.tolist() == pos and [bytes(buf[r[0]:r[1]]) for r in recs] == dates
    n, so, _, _ = o.find_reader(b"prefix 2024-01-15 middle 2024-02-20 suffix")
    assert so.tolist() == [7, 25]
    assert o.find_reader(b"x" * 10000)[0] == 0
    assert o.stream_config(0, 0) == (65536, 1024)


def test_golden_blobs_equal_the_front_end_s_where_both_exist():
    """Not a pin of the oracle (that is the tests above) but of the claim that the two blob sources agree on everything an
    executor reads: for every golden, the oracle gives the same answers on the product front-end's blob."""
    from helpers import compile_blob
    rng = np.random.default_rng(5)
    for e in GOLDENS:
        og, of = Oracle(golden_blob.blob_from_golden(e)), Oracle(compile_blob(e["pattern"]))
        ent = CURATED.get(e["file"])
        inputs = [c["input"].encode("utf-8") for c in ent["cases"]] if ent else [b"2024-01-15 x 10.0.0.1 a@b.cd 12 2024-1-1"]
        inputs += [bytes(rng.choice(np.frombuffer(b"ab01.-:/@ ht", dtype=np.uint8), size=int(rng.integers(1, 60)))) for _ in range(40)]
        for b in inputs:
            assert og.match(b) == of.match(b), (e["file"], b)
            if e["find"]:
                assert og.find(b) == of.find(b), (e["file"], b)
                assert og.find_all(b)[1].tolist() == of.find_all(b)[1].tolist(), (e["file"], b)
