"""Oracle pin #2: on the reference's own corpus (tests/e2e/testdata.json, 238 entries) and on the
input literals of its generated per-pattern tests, the oracle must satisfy the assertions the
reference's generated tests make (internal/compiler/test_gen.go:72-239): MatchBytes, FindBytes and
FindAllBytes(-1) agree with leftmost-first semantics on match text and group texts.  Also checks the
engine labels regengo.Analyze reports (tests/e2e/e2e_test.go:66-95 asserts exactly these).
Expected values: tests/golden/corpus_expected.json (made by tests/golden/make_corpus_fixture.py)."""
import json
import os

import pytest

from helpers import ROOT, compile_blob, compile_json
from oracle import Oracle

with open(os.path.join(ROOT, "tests", "golden", "corpus_expected.json")) as _fh:
    CORPUS = json.load(_fh)

# (every corpus pattern compiles: the \p{...} classes got their Unicode 15.0.0 tables in round 2, tools/gen_unicode_tables.py)
UNSUPPORTED = set()
ENTRIES = [("e2e", i) for i in range(len(CORPUS["e2e"]))] + [("curated", i) for i in range(len(CORPUS["curated"]))]


def texts(b, rec):
    return [(b[rec[2 * g]:rec[2 * g + 1]].decode("utf-8", "replace") if rec[2 * g] >= 0 else "") for g in range(len(rec) // 2)]


def test_corpus_size():
    assert len(CORPUS["e2e"]) == 238
    assert sum(len(e["cases"]) for e in CORPUS["e2e"]) == 1242


@pytest.mark.parametrize("kind,idx", ENTRIES)
def test_oracle_on_corpus_entry(kind, idx):
    ent = CORPUS[kind][idx]
    pat = ent["pattern"]
    if pat in UNSUPPORTED:
        with pytest.raises(ValueError):
            compile_json(pat)
        return
    j = compile_json(pat)
    if "engine_labels" in ent:
        assert sorted(j["engine_labels"]) == sorted(ent["engine_labels"]), pat
    o = Oracle(compile_blob(pat))
    assert (j["find_engine"] != 0) == (ent["n_groups"] > 0)
    for c in ent["cases"]:
        b = c["input"].encode("utf-8")
        assert o.match(b) == c["match"], (pat, c["input"])
        if ent["n_groups"] > 0:
            r = o.find(b)
            assert (None if r is None else texts(b, r)) == c["find"], (pat, c["input"])
            n, recs = o.find_all(b)
            assert [texts(b, list(x)) for x in recs] == c["findall"], (pat, c["input"])
            assert n == len(c["findall"])
