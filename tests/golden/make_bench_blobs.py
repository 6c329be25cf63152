#!/usr/bin/env python3
"""Writes tests/golden/bench_blobs.npz: the program blobs of the bench workloads, so that `bench.py --impl reference`
(the CPU oracle arm) and bench.py's cpu_baseline leg never load the product library at run time.

    python tests/golden/make_bench_blobs.py

The blobs are produced by the product's front-end (rgx_compile -> rgx_program_blob) at commit time;
tests/test_bench_contract.py checks that the committed file still equals what the front-end produces, and the
front-end itself is pinned cell by cell against the reference's generated Go files (tests/test_frontend_goldens.py).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def build_blobs():
    import bench
    import regengo_b200 as rg
    from regengo_b200 import synth
    out = {}
    for name, wl in bench.WORKLOADS.items():
        if "blob" in wl:
            out[wl["blob"]] = np.frombuffer(rg.Pattern(getattr(synth, wl["pattern_name"])).blob(), dtype=np.uint8)
    for k, (pat, _) in enumerate(bench.suite_patterns()):
        out["c4_%03d" % k] = np.frombuffer(rg.Pattern(pat).blob(), dtype=np.uint8)
    return out


if __name__ == "__main__":
    blobs = build_blobs()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bench_blobs.npz"), **blobs)
    print(len(blobs), "blobs,", sum(v.size for v in blobs.values()), "bytes")
