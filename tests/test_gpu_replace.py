"""ReplaceAllBytesAppend over batches on the device (kernels_replace.cuh) against the oracle, bit-exact."""
import json
import os

import numpy as np
import pytest

import regengo_b200 as rg
from regengo_b200 import synth
from oracle import Oracle

from helpers import ROOT

pytestmark = pytest.mark.gpu

CORPUS = json.load(open(os.path.join(ROOT, "tests", "golden", "corpus_expected.json")))
TEMPLATES = ["", "X", "[$0]", "<$1|$2>", "$$ $0 $$", "${1}-${2}-$3", "$9$0", "a$nobody$1b"]


def check(p, o, inputs, template):
    data, offs = rg.pack_inputs(inputs)
    got, goffs = p.replace_all_batch(data, template, offsets=offs)
    exp, eoffs = o.replace_batch(data, offs, template)
    assert np.array_equal(goffs, eoffs), (template, np.nonzero(goffs != eoffs)[0][:5])
    assert np.array_equal(got, exp), template


def test_reference_known_answers_on_device():
    # tests/integration/replace_test.go:37-75,122-134,510-600
    p = rg.Pattern(synth.EMAIL_PATTERN)
    inp = b"Contact alice@example.com and bob@test.org"
    assert p.replace_all(inp, "REDACTED") == b"Contact REDACTED and REDACTED"
    assert p.replace_all(inp, "[$0]") == b"Contact [alice@example.com] and [bob@test.org]"
    assert p.replace_all(inp, "$1@REDACTED.$3") == b"Contact alice@REDACTED.com and bob@REDACTED.org"
    assert p.replace_all(inp, "$user@hidden.$tld") == b"Contact alice@hidden.com and bob@hidden.org"
    assert p.replace_all(inp, "$$user=$user") == b"Contact $user=alice and $user=bob"
    assert p.replace_all(b"no emails here", "[$0]") == b"no emails here"
    assert p.replace_all(b"test@example.com", "[$user]") == b"[test]"
    assert p.replace_all(b"", "$0") == b""
    assert p.replace_all(b"a@b.c d@e.f", "X") == b"X X"
    assert p.replace_all(b"user@example.com " * 1000, "X") == b"X " * 1000
    with pytest.raises(rg.RegengoError):
        p.replace_all(inp, "${unclosed")
    with pytest.raises(rg.RegengoError):
        rg.Pattern(r"\d+").replace_all(b"12", "x")        # no capture groups: Replace* is not generated
    assert rg.launches() > 0


def test_corpus_patterns_replace_batch():
    n_pat = 0
    for ent in CORPUS["e2e"] + CORPUS["curated"]:
        if ent["n_groups"] == 0:
            continue
        p = rg.Pattern(ent["pattern"])
        if p.info.find_engine == 0:
            continue
        o = Oracle(p.blob())
        inputs = [c["input"].encode("utf-8") for c in ent["cases"]]
        inputs += synth.mutate_inputs([c["input"] for c in ent["cases"]], 120, stream=n_pat)
        inputs += [b"", b"a", b"\n", b" ".join(inputs[:6])]
        for t in TEMPLATES[n_pat % 3::3] + ["[$0]"]:
            check(p, o, inputs, t)
        n_pat += 1
    assert n_pat >= 60


def test_empty_matches_relocated_text_and_engines():
    # nullable patterns (empty matches advance one byte and fire once more at the end), the text located by bytes.Index
    # (Q16) with the captures of the true match, skip-restart (Q1), anchors re-applied to every slice (Q3), TDFA engine
    cases = [
        (r"(?P<d>\d*)", {}, [b"ab12c", b"", b"7", b"007 x 9"], "<$d>"),
        (r"(?P<y>\d{4})-(?P<m>\d{2})", {}, [b"12345-67 2024-01", b"12024-01 2024-01", b"2024-012024-01"], "[$y/$m]"),
        (r"^(?P<w>\w+)", {}, [b"ab cd ef", b" ab", b"abc"], "<$w>"),
        (r"(?P<w>\w+)$", {}, [b"ab cd ef", b"ab ", b"abc"], "<$w>"),
        (synth.URL_PATTERN, {}, [b"see http://a.b:80/x and https://c.d/e/f?g", b"http://", b"httphttp://x.y"], "<$protocol|$host|$port|$path>"),
        (synth.URL_PATTERN, {"force_tdfa": True}, [b"see http://a.b:80/x and https://c.d/e/f?g"], "${host}:${port}"),
        (r"(a|ab)(c|bcd)(d*)", {}, [b"abcd abcdd acd", b"xabcdx"], "$3.$2.$1"),
        (r"(?P<k>\w+)=(?P<v>\w*)", {}, [b"a=1 b= c=3", b"=", b"k=v"], "$v=$k"),
    ]
    for pat, kw, inputs, t in cases:
        p = rg.Pattern(pat, **kw)
        check(p, Oracle(p.blob()), inputs, t)


def test_capacity_protocol_and_device_pointers():
    import ctypes as C
    import torch
    from regengo_b200 import _lib
    L = _lib.load()
    p = rg.Pattern(synth.EMAIL_PATTERN)
    o = Oracle(p.blob())
    inputs = [b"mail bob@site.org now", b"", b"x@y.z", b"nothing"] * 500
    data, offs = rg.pack_inputs(inputs)
    exp, eoffs = o.replace_batch(data, offs, "<$user at $domain>")
    dev = torch.device("cuda", 0)
    d_data = torch.from_numpy(data.copy()).to(dev)
    d_offs = torch.from_numpy(offs.view(np.int64)).to(dev)
    n = len(inputs)
    d_out_offs = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    total = C.c_uint64()
    t = b"<$user at $domain>"
    # too small: nothing written, the offsets and the total are valid
    d_small = torch.zeros(16, dtype=torch.uint8, device=dev)
    rc = L.rgx_replace_batch_dev(rg.context(0), p._h, t, len(t), d_data.data_ptr(), d_offs.data_ptr(), n, d_small.data_ptr(), 16,
                                 d_out_offs.data_ptr(), C.byref(total))
    assert rc == _lib.RGX_ECAPACITY and total.value == exp.size
    assert np.array_equal(d_out_offs.cpu().numpy().view(np.uint64), eoffs) and int(d_small.sum()) == 0
    d_out = torch.empty(int(total.value), dtype=torch.uint8, device=dev)
    _lib.check(L.rgx_replace_batch_dev(rg.context(0), p._h, t, len(t), d_data.data_ptr(), d_offs.data_ptr(), n, d_out.data_ptr(),
                                       int(total.value), d_out_offs.data_ptr(), C.byref(total)))
    assert np.array_equal(d_out.cpu().numpy(), exp)
    # empty batch
    got, goffs = p.replace_all_batch([], "x")
    assert got.size == 0 and goffs.tolist() == [0]


def test_large_batch_log_lines():
    # 200 k log lines (the c2 log generator cut at newlines), every email redacted
    p = rg.Pattern(synth.EMAIL_PATTERN)
    o = Oracle(p.blob())
    buf = bytes(synth.make_buffer("log", 8 << 20))
    lines = buf.split(b"\n")[:200000]
    check(p, o, lines, "$user@REDACTED.$tld")
