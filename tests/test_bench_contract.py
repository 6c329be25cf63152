"""bench.py's reference arm runs without a GPU: check the JSON line it prints against the driver's contract."""
import json
import os
import subprocess
import sys

from helpers import ROOT


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                          "--ref-mib", "1", "--ref-seconds", "0.2"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GB/s" and line["higher_is_better"] is True
    assert line["metric"] == "GB/s input scanned for FindAllBytes" and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 3
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["vs_baseline"] is None and line["dtype"] == "u8"
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["config"]["bytes_per_gpu"] == 4 << 30      # what defines the workload, identical on both arms
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb and cb["one_core_value"] > 0
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_other_workloads_and_no_product_library():
    """c4 / c5 on the CPU arm; the arm must not load the product's shared library (it reads committed blobs)."""
    probe = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--workload', %r, '--steps', '1', '--ref-mib', '1', "
             "'--ref-seconds', '0.2']; runpy.run_path(%r, run_name='__main__'); "
             "maps = open('/proc/self/maps').read(); assert 'libregengo_b200' not in maps, 'product library loaded'")
    for wl, metric in (("c4", "GB/s input scanned for batched MatchBytes"), ("c5", "GB/s input scanned for stream.FindReader")):
        out = subprocess.run([sys.executable, "-c", probe % (wl, os.path.join(ROOT, "bench.py"))], capture_output=True, text=True,
                             timeout=600, cwd=ROOT)
        assert out.returncode == 0, out.stderr[-2000:]
        line = json.loads(out.stdout.strip().splitlines()[-1])
        assert line["impl"] == "reference" and line["metric"] == metric and line["value"] > 0


def test_committed_bench_blobs_equal_the_front_end():
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_bench_blobs
    fresh = make_bench_blobs.build_blobs()
    with np.load(os.path.join(ROOT, "tests", "golden", "bench_blobs.npz")) as z:
        assert sorted(z.files) == sorted(fresh)
        for k in fresh:
            assert np.array_equal(z[k], fresh[k]), k
