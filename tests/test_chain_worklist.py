"""The worklist-driven cursor replay (csrc/kernels_chain.cuh, findall_chain4_kernel), restated in Python: parts are
replayed from guessed entries, then only the parts whose predecessor's exit moved are replayed again, exits updated
in place, in ANY order within a pass.  When a worklist comes out empty the result must be the sequential replay."""
import numpy as np


def _replay(records, cursor, total_len, rule):
    """records: sorted (start, len).  Returns (exit cursor, [(start, len, reps)])."""
    out = []
    for s, L in records:
        if s >= cursor and cursor < total_len:
            if rule == "tdfa":       # offset += len(match) from the slice start (compiler.go:630-636)
                step = max(L, 1)
                k = (s - cursor) // step + 1
                cursor += k * step
                out.append((s, L, k))
            else:                    # searchStart = match end, or +1 for an empty match (find.go:452-457)
                cursor = s + L if L else s + 1
                out.append((s, L, 1))
    return cursor, out


def _worklist_resolve(parts, part_pos, total_len, rule, rng):
    n = len(parts)
    entry_used = [0] + [part_pos[p] for p in range(1, n)]
    res = [_replay(parts[p], entry_used[p], total_len, rule) for p in range(n)]          # pass 0
    exits = [r[0] for r in res]
    work = [p for p in range(1, n) if exits[p - 1] != entry_used[p]]                       # seed
    passes = 0
    while work:
        passes += 1
        assert passes <= n + 2
        nxt = []
        for p in rng.permutation(work):          # any order: a pass reads whatever exit its neighbour has right now
            p = int(p)
            entry = exits[p - 1]
            if entry == entry_used[p]:
                continue
            entry_used[p] = entry
            res[p] = _replay(parts[p], entry, total_len, rule)
            if res[p][0] != exits[p]:
                exits[p] = res[p][0]
                if p + 1 < n:
                    nxt.append(p + 1)
        work = nxt
    return [x for r in res for x in r[1]], exits[-1], passes


def test_worklist_replay_equals_sequential_replay():
    rng = np.random.default_rng(5)
    for rule in ("tdfa", "bt"):
        for trial in range(40):
            total_len = int(rng.integers(2000, 40000))
            part_bytes = int(rng.choice([256, 512, 1024]))
            starts = np.unique(rng.integers(0, total_len, size=int(rng.integers(5, 600))))
            lens = rng.integers(0 if rule == "bt" else 1, 60, size=starts.size)
            recs = [(int(s), int(L)) for s, L in zip(starts, lens)]
            n_parts = (total_len + part_bytes - 1) // part_bytes
            parts = [[r for r in recs if p * part_bytes <= r[0] < (p + 1) * part_bytes] for p in range(n_parts)]
            part_pos = [p * part_bytes for p in range(n_parts)]
            want_exit, want = _replay(recs, 0, total_len, rule)
            got, got_exit, passes = _worklist_resolve(parts, part_pos, total_len, rule, rng)
            assert got == want and got_exit == want_exit, (rule, trial, passes)
