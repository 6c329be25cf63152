"""The sharded paths on real hardware: two ranks under NCCL (skipped with fewer than two GPUs).  Each case launches
bench.py with torchrun at an oracle-friendly size; every rank checks ITS shard of the result against the CPU oracle
in-run (bench.py `parity`: the window covers the whole shard at these sizes) and rank 0 reports the conjunction.
Covers SURVEY 8e: one logical FindAll buffer with pre-halo cursor confirmation (c3), input-sharded batched MatchBytes
with the NCCL flag gather (c4), chunk-sharded FindReader (c5)."""
import json
import os
import socket
import subprocess
import sys

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--gpus", str(world), "--steps", "2", "--warmup", "3", "--no-cpu"] + extra
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("extra", [
    ["--workload", "c3", "--gib", "0.046875"],           # 48 MiB per rank: 96 MiB logical buffer, pre-halo + halo of 1 MiB
    ["--workload", "c4", "--inputs", "600000"],
    ["--workload", "c5", "--gib", "0.0625"],
    ["--workload", "c5", "--gib", "0.0625", "--buffer-size", "1048576"],
])
def test_two_ranks_nccl_against_the_oracle(extra):
    line = _run(2, extra)
    assert line["n_gpus"] == 2 and line["gpu_launches"] > 0
    par = line["parity"]
    assert par["ok"] is True and par["ranks_checked"] == 2, par
    assert len(par["result_hash"]) == 2
    assert line["e2e"]["value"] > 0 and line["value"] > 0
