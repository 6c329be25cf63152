"""The scan6 walk image (csrc/device_program.cu) against the reference's TDFA walk (no GPU).

The FindAll scan kernel does not run the reference's loop literally: it takes CHEAP steps (row lookups), logs
EVENTS, and reads match end and tags off the log afterwards (kernels_scan6.cuh, "Exactness of reading the match off
the log").  This test executes that procedure over the image the library actually packs -- rows, descriptors and
tag lists are read from rgx_program_device_image -- and compares, for every start position of every corpus input
of every TDFA pattern on the fast path, with a literal restatement of tdfa.go:929-983 over the TDFA tables of
rgx_program_json.  What it pins is the packing and the claim that the log determines the reference's result; the
kernel's own arithmetic is covered by the GPU parity tests."""
import json
import os

import numpy as np

import regengo_b200 as rg
from regengo_b200 import synth

from helpers import ROOT, compile_json

DEAD, EVMIN, ACC, ACC_EOT = 0xFFC00000, 0x00400000, 1 << 20, 1 << 21


def reference_walk(t, data, start):
    """One start position of findBytesInternal (tdfa.go:929-983) -> (match_end, matchTags) or None."""
    ns, nt = t["num_states"], t["num_tags"]
    tags = [-1] * nt
    tags[0] = start
    state = t["start_begin"] if start == 0 else t["start_any"]
    for tg in (t["init_tags_begin"] if start == 0 else t["init_tags_any"]):
        tags[tg] = start
    match_end, match_tags = -1, None
    if t["accept"][state] or (start == len(data) and t["accept_eot"][state]):
        match_end, match_tags = start, list(tags)
    for i in range(start, len(data)):
        c = data[i]
        if c >= 128:
            break
        nx = t["trans"][state * 128 + c]
        if nx < 0:
            break
        for tg, off in t["actions"].get(str(state * 128 + c), []):
            tags[tg] = i + 1 - off
        state = nx
        if t["accept"][state] or (i == len(data) - 1 and t["accept_eot"][state]):
            for tg, off in t["accept_actions"][state]:
                tags[tg] = i + 1 - off
            match_end, match_tags = i + 1, list(tags)
    if match_end < 0:
        return None
    match_tags[1] = match_end
    return match_end, match_tags


class Image:
    def __init__(self, pattern, **kw):
        p = rg.Pattern(pattern, **kw)
        self.plan = p.device_plan()
        w = p.device_image()
        d = self.plan
        self.w = w[d["w6_off"]: d["w6_off"] + d["scan6_image_bytes"] // 4]
        self.nt = d["tdfa_tags"]

    def entries(self, idx):
        """Tag entries of descriptor idx: [(tag, offset, is_accept_action)]."""
        d = self.plan
        n = int(self.w[d["w6_desc"] + 8 * idx]) >> 24
        out = []
        for i in range(n):
            en = int(self.w[d["w6_fent"] + 8 * idx + i])
            assert (en & 0xFFFF) % 128 == 0
            out.append(((en & 0xFFFF) // 128, (en >> 16) & 0xFF, bool(en >> 31)))
        return out

    def walk(self, data, start):
        """The kernel's procedure: cheap steps, event log, interpretation of the log."""
        d = self.plan
        w = self.w
        row = d["tdfa_start_any"] * 1024
        ri, log, lastacc, eob = start, [], 0, False
        while True:
            if ri >= len(data):
                eob = True
                break
            cell = int(w[row // 4 + data[ri]])
            if cell >= DEAD:
                break
            if cell >= EVMIN:
                log.append((cell >> 22, ri - start))
            row = cell & 0x3FFFFF
            ri += 1
        # the last event whose next state accepts
        for e, (idx, _) in enumerate(log):
            dy = int(w[d["w6_desc"] + 8 * idx])
            if dy & ACC or (e + 1 == len(log) and eob and dy & ACC_EOT):
                lastacc = e + 1
        end_rel = ri - start
        if not lastacc:
            return None, len(log)
        match_end = log[lastacc][1] if lastacc < len(log) else end_rel
        tags = [-1] * self.nt
        tags[0] = 0
        for j in range(d["tdfa_n_init_any"]):
            tags[int(w[d["w6_init"] + j])] = 0
        for e in range(lastacc):
            idx, pos = log[e]
            dy = int(w[d["w6_desc"] + 8 * idx])
            run_end = log[e + 1][1] if e + 1 < len(log) else end_rel
            accp = bool(dy & ACC) or e + 1 == lastacc
            for tg, off, is_acc in self.entries(idx):
                if not is_acc:
                    tags[tg] = pos + 1 - off
                elif accp:
                    tags[tg] = run_end - off
        tags[1] = match_end
        return (match_end, tags), len(log)


def check_pattern(pattern, inputs, **kw):
    img = Image(pattern, **kw)
    if not img.plan["fast_tdfa_scan"]:
        return 0
    t = compile_json(pattern, **kw)["tdfa"]
    assert t["start_begin"] == t["start_any"]
    n = 0
    for data in inputs:
        for start in range(len(data)):
            # the filter's condition is necessary: a start it rejects cannot match
            pre = bytes.fromhex(img.plan["prefix"])[:4]
            tested = [0, 1, 3] if len(pre) == 4 else list(range(len(pre)))   # the prefix bytes the filter compares
            passes = all(start + o < len(data) and data[start + o] == pre[o] for o in tested)
            ref = reference_walk(t, data, start)
            if not passes:
                assert ref is None, (pattern, data, start)
                continue
            got, nlog = img.walk(data, start)
            assert nlog <= img.plan["w6_maxev"], (pattern, data, start, nlog)   # the bound that sizes the kernel's log
            if ref is None:
                assert got is None, (pattern, data, start)
            else:
                assert got is not None, (pattern, data, start)
                rel = [x - start if x >= 0 else -1 for x in ref[1]]
                assert (got[0], got[1]) == (ref[0] - start, rel), (pattern, data, start, got, ref)
            n += 1
    return n


def test_url_pattern_log_interpretation():
    pool, nw, nu = synth._url_pool()
    urls = [bytes(pool.mat[i, :pool.lens[i]]) for i in range(nw, nw + 300)]
    inputs = [b"see " + u + b"and " + u for u in urls[:150]] + [u.rstrip() for u in urls[150:]]   # also URLs cut by the end of the input
    inputs += [b"http://a.b:80/x", b"https://", b"http://a", b"http://a.b:/x", b"http://a.b:8", b"http://\xc3\xa9.com", b"httphttp://x.y/http://z"]
    assert check_pattern(synth.URL_PATTERN, inputs) > 400


def test_first_byte_set_patterns_log_interpretation():
    """Patterns without a literal first byte (the filter is a first-byte range test, every set byte is walked)."""
    semver = r"(?P<major>\d+)\.(?P<minor>\d+)\.(?P<patch>\d+)(?:-(?P<prerelease>[\w.-]+))?(?:\+(?P<build>[\w.-]+))?"
    ipv4 = r"(?P<ip>(?P<a>\d{1,3})\.(?P<b>\d{1,3})\.(?P<c>\d{1,3})\.(?P<d>\d{1,3}))(?::(?P<port>\d{1,5}))?"
    assert check_pattern(semver, [b"v1.2.3 and 10.20.30-rc.1+build.5, 7.8 2024.1.15", b"1.2.3", b"1.2.", b"0.0.0-"], force_tdfa=True) > 40
    assert check_pattern(ipv4, [b"10.0.0.1 192.168.1.254:8080 1.2.3 999.1.1.1:123456", b"1.1.1.1:", b"1234.1.1.1"], force_tdfa=True) > 40
    img = Image(ipv4, force_tdfa=True)
    assert img.plan["prefix_len"] == 0 and 11 < img.plan["w6_maxev"] <= 15      # takes the 15-event log
    assert Image(synth.URL_PATTERN).plan["w6_maxev"] <= 11


def test_corpus_tdfa_patterns_log_interpretation():
    corpus = json.load(open(os.path.join(ROOT, "tests", "golden", "corpus_expected.json")))
    n_pat = 0
    for ent in corpus["e2e"] + corpus["curated"]:
        if ent["n_groups"] == 0:
            continue
        for kw in ({}, {"force_tdfa": True}):
            try:
                inputs = [c["input"].encode("utf-8") for c in ent["cases"]]
                inputs.append(b" ".join(inputs))
                if check_pattern(ent["pattern"], inputs, **kw):
                    n_pat += 1
            except rg.RegengoError:
                pass
    assert n_pat >= 15
