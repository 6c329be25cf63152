"""The fast scan's SWAR prefix filter (csrc/kernels_scan5.cuh, FILTER), restated in numpy over 32-bit words:
z = (w ^ p0) | (w>>8 ^ p1) | (w>>16 ^ p2) | (w>>24 ^ p3) across word boundaries has a zero byte exactly where the
prefix starts; the cheap zero-byte test may only flag an extra byte directly above a true zero byte, which the
"two hits touch" check catches; flags are packed as bit 8*byte + word and decoded as 4*(b & 7) + (b >> 3)."""
import numpy as np


def _splat(b):
    return np.uint32(b) * np.uint32(0x01010101)


def _funnel_r(lo, hi, sh):
    return ((lo.astype(np.uint64) | (hi.astype(np.uint64) << np.uint64(32))) >> np.uint64(sh)).astype(np.uint32)


def _z_words(words, nxt, prefix):
    z = words ^ _splat(prefix[0])
    for t in range(1, len(prefix)):
        z = z | (_funnel_r(words, nxt, 8 * t) ^ _splat(prefix[t]))
    return z


def _hits_naive(data, prefix):
    n = len(data)
    out = np.zeros(n, dtype=bool)
    for i in range(n - len(prefix) + 1):
        out[i] = bytes(data[i:i + len(prefix)]) == prefix
    return out


def test_swar_prefix_filter_matches_naive_search():
    rng = np.random.default_rng(3)
    for prefix in (b"ht", b"id:", b"http", b"aaaa", b"abab", b"\x01\x00"):
        alphabet = np.frombuffer(bytes(set(prefix)) + b"xyz \x00\x01\x02", dtype=np.uint8)
        for trial in range(30):
            n_words = 64
            data = alphabet[rng.integers(0, alphabet.size, size=4 * n_words + 4)].copy()
            for _ in range(6):   # plant occurrences, some overlapping / touching
                p = int(rng.integers(0, data.size - len(prefix)))
                data[p:p + len(prefix)] = np.frombuffer(prefix, dtype=np.uint8)
            words = data[: 4 * n_words + 4].view("<u4")
            w, nxt = words[:n_words], words[1:n_words + 1]
            z = _z_words(w, nxt, prefix)
            cheap = (z - np.uint32(0x01010101)) & ~z & np.uint32(0x80808080)
            exact = ~(((z & np.uint32(0x7F7F7F7F)) + np.uint32(0x7F7F7F7F)) | z | np.uint32(0x7F7F7F7F))
            want = _hits_naive(data, prefix)[: 4 * n_words]
            got_exact = np.zeros(4 * n_words, dtype=bool)
            for j in range(n_words):
                for b in range(4):
                    got_exact[4 * j + b] = bool((int(exact[j]) >> (8 * b + 7)) & 1)
            assert np.array_equal(got_exact, want), prefix
            # the cheap test: a superset, and every extra flag sits directly above a flagged byte in the same word
            extra = cheap & ~exact
            assert not np.any(exact & ~cheap)
            assert np.all((extra & ~(cheap << np.uint32(8))) == 0)
            # packing (groups of 8 words -> bit 8*byte + word) and the kernel's "two hits touch" test
            for g in range(0, n_words, 8):
                e_cheap = np.uint32(0)
                e_exact = np.uint32(0)
                for j in range(8):
                    e_cheap |= cheap[g + j] >> np.uint32(7 - j)
                    e_exact |= exact[g + j] >> np.uint32(7 - j)
                if int(e_cheap & (e_cheap >> np.uint32(8))) == 0:
                    assert e_cheap == e_exact          # no touching hits: the cheap flags are the exact ones
                pos = [4 * (b & 7) + (b >> 3) for b in range(32) if (int(e_exact) >> b) & 1]
                assert sorted(pos) == [int(i) for i in np.nonzero(want[4 * g: 4 * g + 32])[0]]
