"""Host-side planning of the device paths (no GPU): which scan / evaluator / FindReader form a pattern gets.
Pins the analyses in csrc/device_program.cu that the GPU parity tests only exercise implicitly."""
import json
import os

import regengo_b200 as rg
from regengo_b200 import synth

from helpers import ROOT


def test_bench_patterns_take_the_fast_paths():
    url = rg.Pattern(synth.URL_PATTERN).device_plan()
    assert url["find_engine"] == 2 and url["fast_tdfa_scan"] == 1 and url["parallel_findall"] == 1
    # 13 rows of 256 cells, 7 event descriptors + the empty one
    assert bytes.fromhex(url["prefix"]) == b"http" and url["scan6_descriptors"] == 8
    assert 13 * 1024 <= url["scan6_image_bytes"] <= 16 * 1024
    email = rg.Pattern(synth.EMAIL_PATTERN).device_plan()
    assert email["find_engine"] == 1 and email["run_anchor"] == 1 and email["run_literal"] == ord("@")
    # cap3 '@' cap4 \w \w* cap5 '.' cap6 \w \w* cap7 match
    assert email["run_linear_elements"] == 12
    date = rg.Pattern(synth.DATE_CAPTURE_PATTERN).device_plan()
    assert date["straight_line_steps"] == 10 and date["straight_line_classes"] == 2 and date["n_alt"] == 0


def test_shapes_that_must_not_be_linearised():
    # loop followed by a byte INSIDE its class: a shorter run can be needed -> no straight-line continuation
    p = rg.Pattern(r"(\w+)@(\w+)x(\w+)").device_plan()
    assert p["run_anchor"] == 1 and p["run_linear_elements"] == 0
    # alternation after the anchor
    p = rg.Pattern(r"(\w+)@(a|bc)").device_plan()
    assert p["run_anchor"] == 1 and p["run_linear_elements"] == 0
    # literal inside the leading class: not a run anchor at all
    assert rg.Pattern(r"(\w+)a(\d)").device_plan()["run_anchor"] == 0
    # optional group / loop: not a straight-line program
    assert rg.Pattern(r"(\d{2})-?(\d)").device_plan()["straight_line_steps"] == 0
    assert rg.Pattern(r"(\d+)").device_plan()["straight_line_steps"] == 0
    # anchored and nullable patterns stay on the sequential / generic paths
    assert rg.Pattern(r"^(\w+)").device_plan()["parallel_findall"] == 0
    assert rg.Pattern(r"(a*)").device_plan()["nullable"] == 1


def test_plan_invariants_over_the_corpus():
    corpus = json.load(open(os.path.join(ROOT, "tests", "golden", "corpus_expected.json")))
    n = 0
    for ent in corpus["e2e"] + corpus["curated"]:
        try:
            p = rg.Pattern(ent["pattern"])
        except rg.RegengoError:
            continue
        d = p.device_plan()
        n += 1
        if d["fast_tdfa_scan"]:
            assert d["find_engine"] == 2 and (d["prefix_len"] >= 1 or d["gen_kind"] == 1) and d["nullable"] == 0
            assert 0 < d["scan6_image_bytes"] <= 96 * 1024 and 1 <= d["scan6_descriptors"] <= 1023
        if d["run_linear_elements"]:
            assert d["run_anchor"] == 1 and d["find_engine"] == 1
        if d["straight_line_steps"]:
            assert d["n_alt"] == 0 and d["n_empty"] == 0 and 1 <= d["straight_line_classes"] <= 8 and d["straight_line_steps"] <= 32
    assert n >= 230
