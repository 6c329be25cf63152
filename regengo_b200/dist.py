"""Multi-GPU sharding of the matching path (one process per GPU, torch.distributed for the plumbing).

The three paths shard differently:

  * batched MatchBytes / FindBytes -- inputs are independent: contiguous index ranges balanced by
    bytes, no data-path collective; results are gathered with one collective (`gather_flags`).
  * stream.FindReader              -- chunks are independent given (BufferSize, MaxLeftover): contiguous
    chunk ranges, each rank holding its stride-aligned span plus the halo its last chunk reads.
  * FindAllBytes over ONE buffer   -- the scan is independent per shard (plus a halo for matches that
    start in the shard and end in the next), but the reference's cursor runs through the whole buffer,
    so rank r needs rank r-1's exit cursor.  Every rank first replays the cursor from a guessed entry,
    the 8-byte exit cursors are all-gathered, and ranks whose entry was wrong redo the replay (not the
    scan) until nothing changes -- at most world_size rounds, normally two.

Only the exchange logic lives here; it takes the per-rank `resolve(entry) -> (exit, payload)`
callable as a parameter so that it can be tested on CPU with the gloo backend.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class Shard:
    rank: int
    world: int
    start: int       # first byte of the shard in the logical buffer
    shard_len: int   # bytes owned
    buf_len: int     # bytes held (shard + halo)
    is_last: bool


def shard_buffer(total_len, world, rank, halo, align=1 << 20):
    """Equal contiguous shards (multiples of `align`, remainder on the last rank) plus a halo."""
    per = (total_len // world) // align * align
    if per == 0:
        per = -(-total_len // world)
    start = min(rank * per, total_len)
    end = total_len if rank == world - 1 else min(start + per, total_len)
    buf_end = min(end + halo, total_len)
    # is_last: the held bytes end at the end of the logical buffer (a walk reaching it has seen everything)
    return Shard(rank, world, start, end - start, buf_end - start, buf_end >= total_len)


def shard_inputs_by_bytes(offsets, world):
    """Split inputs [0, n) into `world` contiguous index ranges with (nearly) equal byte counts.
    offsets: uint64[n+1].  Returns a list of (first, last_exclusive)."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n = offsets.size - 1
    total = int(offsets[-1] - offsets[0])
    cuts = [0]
    for r in range(1, world):
        target = int(offsets[0]) + total * r // world
        cuts.append(int(np.searchsorted(offsets, target, side="left")))
    cuts.append(n)
    cuts = [min(max(c, 0), n) for c in cuts]
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def shard_chunks(n_chunks, world, rank):
    """Contiguous chunk-index range of FindReader for this rank: (first, count)."""
    per = -(-n_chunks // world)
    first = min(rank * per, n_chunks)
    return first, min(per, n_chunks - first)


def chunk_span(first, count, buffer_size, max_leftover, total_len):
    """Stream bytes [lo, hi) that chunks first..first+count-1 read (their stride-aligned span + halo)."""
    stride = buffer_size - max_leftover
    if count <= 0:
        return 0, 0
    lo = first * stride
    hi = min((first + count - 1) * stride + buffer_size, total_len)
    return min(lo, total_len), hi


def resolve_cursor_chain(resolve, rank, world, shard_start, all_gather_i64, first_guess=None, finish=None, all_starts=None):
    """Make every rank's entry cursor equal its predecessor's exit cursor.

    resolve(entry_global) -> (exit_global, payload): replays this rank's records from `entry_global`
    (a position in the logical buffer) and returns where the cursor leaves the shard.
    finish() -> payload (optional): when given, `resolve` only replays the cursor and `finish` produces
    the output once, after the entries have settled (saves the output pass of every discarded round).
    all_gather_i64(v) -> list of every rank's v (a collective; every rank calls it the same number of
    times).  all_starts (optional): every rank's shard_start; with it each rank can tell from the gathered
    exits alone whether ANY rank still has to redo its replay, so a round costs one collective instead of two.
    Returns (entry, exit, payload, rounds)."""
    entry = 0 if rank == 0 else (shard_start if first_guess is None else first_guess)
    entries = None
    if all_starts is not None and first_guess is None:
        entries = [0] + [int(x) for x in all_starts[1:]]
    rounds = 0
    exit_cur, payload = resolve(entry)
    while True:
        rounds += 1
        exits = all_gather_i64(exit_cur)
        want = 0 if rank == 0 else int(exits[rank - 1])
        changed = int(want != entry)
        if entries is not None:
            wants = [0] + [int(exits[r - 1]) for r in range(1, world)]
            any_changed = int(any(w != e for w, e in zip(wants, entries)))
            entries = wants
        else:
            any_changed = max(all_gather_i64(changed))
        if not any_changed:
            if finish is not None:
                payload = finish()
            return entry, exit_cur, payload, rounds
        if changed:
            entry = want
            exit_cur, payload = resolve(entry)
        if rounds > world + 1:
            raise RuntimeError("cursor exchange did not converge")


def settle_pre_halo_chain(run_pre, run_from, rank, world, shard_start, all_gather_pair):
    """FindAll over one buffer with PRE-HALOS: every rank > 0 replays the cursor through the last stretch of its
    predecessor's shard (from a guess) before its own, so the cursor it carries into its shard is almost surely the true
    one and ONE all-gather per step only has to confirm it.

    run_pre() -> (entry_global, exit_global, payload): scan + replay + output with the pre-halo (rank 0: from cursor 0).
    run_from(entry_global) -> (exit_global, payload): the same from an explicit entry cursor (the redo of a rank whose
    carried cursor was wrong; its new exit is re-checked against its successor).
    all_gather_pair((a, b)) -> list of every rank's (a, b); every rank calls it the same number of times.
    Returns (entry_global, exit_global, payload, rounds)."""
    entry, exit_cur, payload = run_pre()
    rounds = 0
    while True:
        rounds += 1
        pairs = all_gather_pair((entry, exit_cur))
        bad = pre_halo_bad_ranks(pairs, world)
        if not bad:
            return entry, exit_cur, payload, rounds
        if rank in bad:
            entry = int(pairs[rank - 1][1])
            exit_cur, payload = run_from(entry)
        if rounds > world + 1:
            raise RuntimeError("cursor exchange did not converge")


def torch_all_gather_pair(group=None, device=None):
    """all_gather of two int64 per rank (NCCL on GPUs, gloo on CPU): one collective, one device-to-host read."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    src = torch.zeros(2, dtype=torch.int64, device=device)
    dst = torch.zeros(2 * world, dtype=torch.int64, device=device)

    def fn(pair):
        src.copy_(torch.tensor([int(pair[0]), int(pair[1])], dtype=torch.int64), non_blocking=False)
        if hasattr(dist, "all_gather_into_tensor") and (device is not None and str(device).startswith("cuda")):
            dist.all_gather_into_tensor(dst, src, group=group)
            flat = dst.tolist()
        else:
            out = [torch.zeros_like(src) for _ in range(world)]
            dist.all_gather(out, src, group=group)
            flat = [int(v) for t in out for v in t.tolist()]
        return [(flat[2 * r], flat[2 * r + 1]) for r in range(world)]
    return fn


def torch_all_gather_pair_async(group=None, device=None):
    """The same collective, split in two: post(pair) enqueues the all-gather of this rank's two int64 and returns a
    handle; collect(handle) waits for it and returns every rank's pair.  A caller that runs the same sharded step over
    and over posts the confirmation of step k and collects it while step k + 1 computes (bench.py): the collective's
    latency and the ranks' skew stay off the critical path.  Two buffer sets alternate, so at most one handle may be
    outstanding when the next is posted."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    on_gpu = device is not None and str(device).startswith("cuda")
    sets = []
    for _ in range(2):
        pin = torch.zeros(2, dtype=torch.int64)
        if on_gpu:
            pin = pin.pin_memory()
        sets.append((pin, torch.zeros(2, dtype=torch.int64, device=device), torch.zeros(2 * world, dtype=torch.int64, device=device)))
    state = {"k": 0}

    def post(pair):
        pin, src, dst = sets[state["k"] & 1]
        state["k"] += 1
        pin[0] = int(pair[0])
        pin[1] = int(pair[1])
        src.copy_(pin, non_blocking=True)
        if on_gpu and hasattr(dist, "all_gather_into_tensor"):
            work = dist.all_gather_into_tensor(dst, src, group=group, async_op=True)
            return (work, dst, None)
        out = [torch.zeros_like(src) for _ in range(world)]
        work = dist.all_gather(out, src, group=group, async_op=True)
        return (work, None, out)

    def collect(handle):
        work, dst, out = handle
        work.wait()
        flat = dst.tolist() if dst is not None else [int(v) for t in out for v in t.tolist()]
        return [(flat[2 * r], flat[2 * r + 1]) for r in range(world)]
    return post, collect


def pre_halo_bad_ranks(pairs, world):
    """The confirmation rule of settle_pre_halo_chain on gathered (entry, exit) pairs: rank r's carried entry cursor must
    equal rank r - 1's exit cursor (rank 0 enters at 0).  Returns the list of ranks whose entry is wrong."""
    want = [0] + [int(pairs[r - 1][1]) for r in range(1, world)]
    return [r for r in range(world) if int(pairs[r][0]) != want[r]]


def torch_all_gather_i64(group=None, device=None):
    """all_gather of one int64 per rank with torch.distributed (NCCL on GPUs, gloo on CPU): one collective
    into a preallocated tensor, one device-to-host read."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    src = torch.zeros(1, dtype=torch.int64, device=device)
    dst = torch.zeros(world, dtype=torch.int64, device=device)

    def fn(v):
        src.fill_(int(v))
        if hasattr(dist, "all_gather_into_tensor") and (device is not None and str(device).startswith("cuda")):
            dist.all_gather_into_tensor(dst, src, group=group)
            return dst.tolist()
        out = [torch.zeros_like(src) for _ in range(world)]
        dist.all_gather(out, src, group=group)
        return [int(x.item()) for x in out]
    return fn


def gather_flags(local_flags, group=None, dst=0):
    """Gather per-input result bytes (MatchBytes flags) of every rank on `dst`, in rank order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([local_flags.numel()], dtype=torch.int64, device=local_flags.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(x.item()) for x in sizes]
    mx = max(sizes) if sizes else 0
    padded = torch.zeros(mx, dtype=local_flags.dtype, device=local_flags.device)
    padded[: local_flags.numel()] = local_flags
    out = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    if dist.get_rank(group) != dst:
        return None
    return torch.cat([o[:s] for o, s in zip(out, sizes)])
