"""Differential fuzzing of the CUDA path against the oracle: random patterns over the corpus'
grammar -- literals, classes, groups, alternation, greedy / lazy / counted repeats, anchors, word boundaries -- and random
inputs over a small alphabet, through MatchBytes, FindBytes, FindAllBytes and FindReader.  Patterns the front-end rejects
are skipped (the rejection itself is pinned elsewhere); everything it accepts must agree bit for bit."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

pytestmark = pytest.mark.gpu

import regengo_b200 as rg  # noqa: E402
from oracle import Oracle  # noqa: E402

ATOMS = ["a", "b", "c", "1", "2", r"\d", r"\w", r"\s", "[ab]", "[^a]", "[a-c1]", ".", "-", "@", r"\."]


def atom():
    return st.sampled_from(ATOMS)


def quant(inner):
    return st.tuples(inner, st.sampled_from(["", "", "*", "+", "?", "*?", "+?", "{2}", "{1,3}", "{0,2}?"])).map(lambda t: t[0] + t[1])


def group(inner):
    return st.tuples(st.sampled_from(["(", "(?:", "(?P<g>"]), inner).map(lambda t: t[0] + t[1] + ")")


def expr(depth):
    if depth == 0:
        return quant(atom())
    sub = expr(depth - 1)
    seq = st.lists(st.one_of(quant(atom()), quant(group(sub))), min_size=1, max_size=4).map("".join)
    alt = st.lists(seq, min_size=1, max_size=3).map("|".join)
    return alt


def runs_forever_in_the_reference(p):
    """A hole of the reference itself: a pattern flagged for the Thompson engine whose program has more than 64 states
    gets a plain goto-machine WITHOUT memoisation for MatchBytes (compiler.go:262-285 falls back, compiler.go:127-132 has
    already decided `useMemoization = false`), which never terminates on loops over empty matches.  The oracle, being a
    literal restatement, hangs there too; the device reports scratch exhaustion.  Nothing to compare."""
    import json
    j = json.loads(p.json())
    return bool(j["catastrophic_risk"]) and j["match_engine"] == 0 and not j["match_memo"]


PATTERN = st.tuples(st.sampled_from(["", "", "", "^", r"\b"]), expr(2), st.sampled_from(["", "", "", "$", r"\b"])).map(
    lambda t: t[0] + "(" + t[1] + ")" + t[2])    # at least one capture group: Find* exist
INPUT = st.text(alphabet="ab c12-@.\n", min_size=0, max_size=40).map(lambda s: s.encode())


@settings(max_examples=60, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(PATTERN, st.lists(INPUT, min_size=1, max_size=6))
def test_fuzz_match_find_findall(pattern, inputs):
    try:
        p = rg.Pattern(pattern)
    except rg.RegengoError:
        return
    if runs_forever_in_the_reference(p) or p.num_cap > 32:      # (more than 15 groups: rejected by the device engines, loudly)
        return
    o = Oracle(p.blob())
    data, offs = rg.pack_inputs(inputs)
    assert np.array_equal(p.match_batch(data, offs), o.match_batch(data, offs)), pattern
    gf, gr = p.find_batch(data, offs)
    ef, er = o.find_batch(data, offs)
    assert np.array_equal(gf, ef), pattern
    m = ef.astype(bool)
    assert np.array_equal(gr[m], er[m]), pattern
    for b in inputs:
        joined = (b + b" ") * 9
        cnt, recs = p.find_all_offsets(joined)
        ecnt, erecs = o.find_all(joined)
        assert cnt == ecnt and np.array_equal(recs, erecs), (pattern, joined)


@settings(max_examples=12, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(PATTERN, INPUT)
def test_fuzz_find_reader(pattern, chunk):
    try:
        p = rg.Pattern(pattern)
    except rg.RegengoError:
        return
    if runs_forever_in_the_reference(p) or p.num_cap > 32 or p.info.find_memo:
        # (memoised FindBytes clears numInst * (len + 1) visited bits at every restart, compiler.go:812-818: on a 64 KiB
        # chunk that is the reference's own quadratic cost, replayed by one device thread -- minutes, not a parity matter)
        return
    o = Oracle(p.blob())
    data = (chunk + b"\n") * 1700          # one full 64 KiB buffer and a tail
    try:
        en, eso, eci, erecs = o.find_reader(data)
    except ValueError:
        return
    n, so, ci, recs = p.find_reader_offsets(data)
    assert n == en and np.array_equal(so, eso) and np.array_equal(ci, eci) and np.array_equal(recs, erecs), pattern


TEMPLATE = st.lists(st.sampled_from(["$0", "$1", "$2", "$g", "${g}", "${1}", "$$", "$", "<", ">", "x", " ", "$9", "$nobody", "\xe9"]),
                    min_size=0, max_size=5).map("".join)


@settings(max_examples=60, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(PATTERN, TEMPLATE, st.lists(INPUT, min_size=1, max_size=6))
def test_fuzz_replace_all(pattern, template, inputs):
    """ReplaceAllBytesAppend: empty matches, anchors re-applied to every slice, relocated match text, unset groups."""
    try:
        p = rg.Pattern(pattern)
    except rg.RegengoError:
        return
    if runs_forever_in_the_reference(p) or p.num_cap > 32 or p.info.find_memo:
        return
    o = Oracle(p.blob())
    inputs = inputs + [b" ".join(inputs) * 3]
    data, offs = rg.pack_inputs(inputs)
    exp, eoffs = o.replace_batch(data, offs, template)
    got, goffs = p.replace_all_batch(data, template, offsets=offs)
    assert np.array_equal(goffs, eoffs) and np.array_equal(got, exp), (pattern, template)
