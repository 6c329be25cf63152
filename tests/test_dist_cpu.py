"""Host-side multi-GPU logic on CPU: world_size-2 (and 3) gloo groups exercise the cursor exchange of
sharded FindAllBytes, the byte-balanced input split and the FindReader chunk split.  The per-rank
"device" work is played by the oracle so that the test needs no GPU."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import compile_blob
from oracle import Oracle
from regengo_b200 import dist as rd
from regengo_b200 import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _records_for_shard(o, buf, sh):
    """All (start, end) of matching starts inside the shard: what the device scan produces."""
    recs = []
    pos = sh.start
    end = sh.start + sh.shard_len
    while pos < end:
        r = o.find(buf[pos: sh.start + sh.buf_len] if not sh.is_last else buf[pos:])
        if r is None:
            break
        s, e = pos + r[0], pos + r[1]
        if s >= end:
            break
        recs.append((s, e, r))
        pos = s + 1
    return recs


def _replay_tdfa(recs, entry, total_len):
    """The TDFA FindAll cursor rule (compiler.go:618-636) over sparse records, from `entry`."""
    cur = entry
    out = []
    for s, e, r in recs:
        if s >= cur and cur < total_len:
            L = max(e - s, 1)
            k = (s - cur) // L + 1
            cur += k * L
            out.append((s, e, k))
    return cur, out


def _worker(rank, world, port, total_len, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        o = Oracle(compile_blob(synth.URL_PATTERN))
        buf = synth.make_buffer("url", total_len)
        sh = rd.shard_buffer(total_len, world, rank, halo=4096, align=4096)
        recs = _records_for_shard(o, buf, sh)
        gather = rd.torch_all_gather_i64()
        entry, exit_cur, payload, rounds = rd.resolve_cursor_chain(lambda e: _replay_tdfa(recs, e, total_len), rank, world, sh.start, gather)
        # the bench's variant: replay-only rounds, one collective per round (every rank knows all shard starts),
        # output produced once at the end -- must settle on the same cursors and the same records
        last = {}

        def replay_only(e):
            last["exit"], last["out"] = _replay_tdfa(recs, e, total_len)
            return last["exit"], None
        starts = [rd.shard_buffer(total_len, world, r, halo=4096, align=4096).start for r in range(world)]
        entry2, exit2, payload2, rounds2 = rd.resolve_cursor_chain(replay_only, rank, world, sh.start, gather, finish=lambda: last["out"],
                                                                   all_starts=starts)
        assert (entry2, exit2, payload2) == (entry, exit_cur, payload) and rounds2 == rounds
        n_local = sum(k for _, _, k in payload)
        counts = gather(n_local)
        q.put((rank, entry, exit_cur, n_local, [(s, e, k) for s, e, k in payload][:3], rounds, counts))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_findall_cursor_exchange_gloo(world):
    total_len = 6 * 65536 + 123
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total_len, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the reference result on the whole buffer
    o = Oracle(compile_blob(synth.URL_PATTERN))
    buf = synth.make_buffer("url", total_len)
    n, recs = o.find_all(buf)
    assert sum(r[3] for r in res) == n
    assert res[0][6] == [r[3] for r in res]
    # every rank's entry is its predecessor's exit
    for r in range(1, world):
        assert res[r][1] == res[r - 1][2]
    assert res[0][1] == 0


def test_shard_inputs_by_bytes():
    rng = np.random.default_rng(1)
    lens = rng.integers(0, 80, size=1000)
    offs = np.zeros(1001, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    for world in (1, 2, 3, 8):
        parts = rd.shard_inputs_by_bytes(offs, world)
        assert parts[0][0] == 0 and parts[-1][1] == 1000
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        sizes = [int(offs[b] - offs[a]) for a, b in parts]
        assert max(sizes) - min(sizes) <= 160


def test_shard_chunks_and_spans_cover_the_stream():
    B, L, total = 65536, 1024, 1_000_000
    stride = B - L
    full = (total - B) // stride + 1
    n_chunks = full + 1
    for world in (1, 2, 4, 8):
        seen = 0
        for r in range(world):
            first, cnt = rd.shard_chunks(n_chunks, world, r)
            assert first == seen
            seen += cnt
            lo, hi = rd.chunk_span(first, cnt, B, L, total)
            if cnt:
                assert lo == first * stride and hi <= total and hi - lo <= cnt * stride + L
        assert seen == n_chunks


def test_shard_buffer_geometry():
    for world in (1, 2, 4, 8):
        total = (1 << 24) + 777
        cover = 0
        for r in range(world):
            sh = rd.shard_buffer(total, world, r, halo=1 << 16)
            assert sh.start == cover
            cover += sh.shard_len
            assert sh.buf_len >= sh.shard_len and sh.start + sh.buf_len <= total
            assert sh.is_last == (sh.start + sh.buf_len == total)
        assert cover == total


def _worker_pre(rank, world, port, total_len, pre, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        o = Oracle(compile_blob(synth.URL_PATTERN))
        buf = synth.make_buffer("url", total_len)
        sh = rd.shard_buffer(total_len, world, rank, halo=4096, align=4096)
        recs = _records_for_shard(o, buf, sh)
        # the pre-halo: the records of the last `pre` bytes of the predecessor's shard
        pre_start = max(0, sh.start - pre)
        pre_sh = rd.Shard(rank, world, pre_start, sh.start - pre_start, sh.start + 4096 - pre_start, False)
        pre_recs = _records_for_shard(o, buf, pre_sh) if rank > 0 else []
        redone = [0]

        def run_pre():
            if rank == 0:
                ex, out = _replay_tdfa(recs, 0, total_len)
                return 0, ex, out
            carried, _ = _replay_tdfa(pre_recs, pre_start, total_len)      # guess: the cursor stands at the pre-halo's first byte
            ex, out = _replay_tdfa(recs, carried, total_len)
            return carried, ex, out

        def run_from(entry):
            redone[0] += 1
            return _replay_tdfa(recs, entry, total_len)
        entry, exit_cur, payload, rounds = rd.settle_pre_halo_chain(run_pre, run_from, rank, world, sh.start, rd.torch_all_gather_pair())
        q.put((rank, entry, exit_cur, sum(k for _, _, k in payload), rounds, redone[0]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,pre", [(2, 65536), (3, 65536), (3, 300), (3, 10 ** 9)])
def test_pre_halo_cursor_confirmation_gloo(world, pre):
    """The pre-halo protocol of the sharded FindAll (dist.settle_pre_halo_chain): a rank carries the cursor through a
    pre-halo from a guess; cursors merge after some records, so a long pre-halo (the bench uses 1 MiB, about 3500
    records) makes the carried cursor the true one and one gather confirms it.  Short ones (64 KiB, 300 bytes) may
    fail the check: the ranks then redo from the right entry -- the result is the sequential one either way.  A
    pre-halo reaching back to byte 0 starts from the true cursor and is always confirmed at once."""
    total_len = 6 * 65536 + 123
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_pre, args=(r, world, port, total_len, pre, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = Oracle(compile_blob(synth.URL_PATTERN))
    n, _ = o.find_all(synth.make_buffer("url", total_len))
    assert sum(r[3] for r in res) == n
    assert res[0][1] == 0
    for r in range(1, world):
        assert res[r][1] == res[r - 1][2]          # every entry is the predecessor's exit
    if pre >= total_len:
        assert all(r[5] == 0 for r in res) and all(r[4] == 1 for r in res)   # confirmed at once: no redo, one round


def _worker_async_pairs(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        post, collect = rd.torch_all_gather_pair_async()
        # the confirmation of step k is posted after step k and collected during step k + 1 (bench.py's sharded FindAll)
        seen = []
        pending = None
        for step in range(5):
            pair = (0 if rank == 0 else 1000 * rank + step, 1000 * (rank + 1) + step)      # entry == predecessor's exit
            if step == 3 and rank == 1:
                pair = (7, pair[1])                                                           # rank 1 carries a wrong cursor once
            h = post(pair)
            if pending is not None:
                seen.append(rd.pre_halo_bad_ranks(collect(pending), world))
            pending = h
        seen.append(rd.pre_halo_bad_ranks(collect(pending), world))
        q.put((rank, seen))
    finally:
        dist.destroy_process_group()


def test_deferred_cursor_confirmation_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_async_pairs, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # every rank sees the same verdict per step: all fine except step 3, where rank 1's entry is wrong
    assert got[0] == got[1] == [[], [], [], [1], []]
    assert rd.pre_halo_bad_ranks([(0, 10), (10, 20), (21, 30)], 3) == [2]
    assert rd.pre_halo_bad_ranks([(5, 10), (10, 20)], 2) == [0]
