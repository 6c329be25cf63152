"""ReplaceAllBytesAppend: the oracle's restatement of replace.go:192-273 and replace/template.go against the
reference's own known answers, and the product's host-side template parser against both.  No GPU."""
import numpy as np
import pytest

import regengo_b200 as rg
from regengo_b200 import synth
from oracle import Oracle

EMAIL = synth.EMAIL_PATTERN          # tests/integration/replace_test.go:18


def pair(pattern, **kw):
    p = rg.Pattern(pattern, **kw)
    return p, Oracle(p.blob())


def test_reference_known_answers():
    # tests/integration/replace_test.go:37-75 (program) and :122-134 (expected output)
    p, o = pair(EMAIL)
    inp = b"Contact alice@example.com and bob@test.org"
    assert o.replace_all(inp, "REDACTED") == b"Contact REDACTED and REDACTED"
    assert o.replace_all(inp, "[$0]") == b"Contact [alice@example.com] and [bob@test.org]"
    assert o.replace_all(inp, "$1@REDACTED.$3") == b"Contact alice@REDACTED.com and bob@REDACTED.org"
    assert o.replace_all(inp, "$user@hidden.$tld") == b"Contact alice@hidden.com and bob@hidden.org"
    assert o.replace_all(inp, "$$user=$user") == b"Contact $user=alice and $user=bob"
    assert o.replace_all(b"no emails here", "[$0]") == b"no emails here"
    assert o.replace_all(b"test@example.com", "[$user]") == b"[test]"
    # :176-226
    assert o.replace_all(b"test@example.com and admin@site.org", "$user@REDACTED.$tld") == b"test@REDACTED.com and admin@REDACTED.org"
    # edge cases, :510-600
    assert o.replace_all(b"", "$0") == b""
    assert o.replace_all(b"a@b.c rest of text", "X") == b"X rest of text"
    assert o.replace_all(b"rest of text a@b.c", "X") == b"rest of text X"
    assert o.replace_all(b"a@b.c d@e.f", "X") == b"X X"
    assert o.replace_all(b"a@b.c", "X") == b"X"
    assert o.replace_all(b"user@example.com " * 1000, "X") == b"X " * 1000
    # :867-869
    assert o.replace_all(b"Contact alice@example.com or bob@test.org", "$user@REDACTED.$tld") == b"Contact alice@REDACTED.com or bob@REDACTED.org"


# replace/template_test.go:14-118: (template, segments) with segments as (kind, value); names resolved against
# a pattern with groups user=1, domain=2, tld=3 the way the generated switch does (unknown: nothing)
PARSE_KATS = [
    ("", []),
    ("hello world", [(-1, b"hello world")]),
    ("$0", [(0, b"")]),
    ("$1", [(1, b"")]),
    ("$12", []),                               # CaptureIndex 12: CaptureByIndex returns nil
    ("$name", []),                             # no group of that name
    ("$$", [(-1, b"$")]),
    ("${1}", [(1, b"")]),
    ("${user}", [(1, b"")]),
    ("${0}", [(0, b"")]),
    ("$user@REDACTED.$tld", [(1, b""), (-1, b"@REDACTED."), (3, b"")]),
    ("$1 by $author ($2)", [(1, b""), (-1, b" by  ("), (2, b""), (-1, b")")]),
    ("cost: $", [(-1, b"cost: $")]),
    ("$ not a ref", [(-1, b"$ not a ref")]),
    ("$2x", [(2, b""), (-1, b"x")]),
    ("$23", []),                               # two digits are one index
    ("$domain_x", []),                         # names are greedy: "domain_x"
    ("${domain}_x", [(2, b""), (-1, b"_x")]),
]
PARSE_ERRORS = ["${unclosed", "${}", "${1abc}", "${123abc}", "x${a b}", "${-}"]


def test_template_parse_known_answers():
    p, o = pair(EMAIL)
    for t, want in PARSE_KATS:
        assert p.template_segments(t) == want, t
    for t in PARSE_ERRORS:
        with pytest.raises(rg.RegengoError) as ei:
            p.template_segments(t)
        assert "invalid replace template: at position" in str(ei.value)
        with pytest.raises(ValueError):
            o.replace_all(b"a@b.c", t)


def expand(segments, data, rec):
    out = b""
    for g, lit in segments:
        if g < 0:
            out += lit
        elif rec[2 * g] >= 0:
            out += data[rec[2 * g]: rec[2 * g + 1]]
    return out


def test_product_parser_agrees_with_oracle_parser():
    """Two independent restatements of replace.Parse (C in the oracle, C++ in the library) expand every template the
    same way -- including the byte-as-rune quirk of the unbraced name scanner (a byte >= 0x80 is judged as a Latin-1
    code point: 0xC3 is a letter, 0xA9 is not)."""
    p, o = pair(EMAIL)
    data = b"bob@site.org"
    found, rec = o.find_batch(np.frombuffer(data, dtype=np.uint8), np.array([0, len(data)], dtype=np.uint64))
    assert found[0]
    rng = np.random.default_rng(5)
    atoms = [b"$", b"$$", b"$0", b"$1", b"$3", b"$4", b"$12", b"$user", b"$tld", b"$nobody", b"${2}", b"${tld}", b"${9}", b"${nobody}",
             b"x", b" ", b"_", b"7", b"\xc3\xa9", b"\xaa", b"\xd7", b"{", b"}", b"-"]
    n = 0
    for _ in range(3000):
        t = b"".join(atoms[int(k)] for k in rng.integers(0, len(atoms), size=int(rng.integers(0, 7))))
        try:
            want = o.replace_all(data, t)
        except ValueError:
            with pytest.raises(rg.RegengoError):
                p.template_segments(t)
            continue
        except NotImplementedError:
            continue
        assert expand(p.template_segments(t), data, rec[0]) == want, t
        n += 1
    assert n > 2000


def test_braced_names_are_checked_rune_by_rune():
    # isValidIdentifier ranges over RUNES (template.go:251-267): a UTF-8 letter is fine (and names no group), a
    # non-letter rune or a broken sequence is "invalid capture name"
    p, _ = pair(EMAIL)
    assert p.template_segments("${été}") == []
    assert p.template_segments("${中9}") == []
    for t in ["${©}", b"${\xc3}", b"${a\xff}", "${9é}"]:
        with pytest.raises(rg.RegengoError):
            p.template_segments(t)


def test_empty_matches_and_relocated_text():
    # empty matches advance one byte and also fire at the end of the input (replace.go:252-261); the match text is
    # located with bytes.Index from the slice start (Q16) while the captures stay those of the true match
    p, o = pair(r"(?P<d>\d*)")
    # (unlike Go's regexp.ReplaceAll, an empty match right behind a match counts: after "12" the loop searches "c" and
    # finds the empty match at its start)
    assert o.replace_all(b"ab12c", "<$d>") == b"<>a<>b<12><>c<>"
    assert o.replace_all(b"", "<$d>") == b"<>"
    assert o.replace_all(b"7", "<$d>") == b"<7><>"      # the loop runs once more on the empty rest
    p2, o2 = pair(r"(?P<y>\d{4})-(?P<m>\d{2})")
    # skip-restart (Q1): the attempt at "12345-67" fails at offset 4 and restarts behind it
    assert o2.replace_all(b"12345-67 2024-01", "[$y/$m]") == b"12345-67 [2024/01]"
    # ... and the copy of the text it jumped over is where bytes.Index reports the match: "12024-01 2024-01" -> the
    # attempt at 0 fails at offset 4, the match is found at 9, its text first occurs at 1
    assert o2.replace_all(b"12024-01 2024-01", "[$y/$m]") == b"1[2024/01] [2024/01]"
