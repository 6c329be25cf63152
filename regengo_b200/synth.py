"""Seeded synthetic workloads for the BASELINE.json configs.

Every generator is block-addressable: block b of a buffer depends only on (seed, b), so any window
of a multi-GiB buffer can be regenerated on the CPU for the oracle without building the whole thing.
numpy only (host); bench.py uploads the blocks to the device before the timed region.
"""
import numpy as np

BLOCK = 1 << 20  # bytes per independently generated block

SEED_C1, SEED_C2, SEED_C3, SEED_C4, SEED_C5 = 0x5EED0001, 0x5EED0002, 0x5EED0003, 0x5EED0004, 0x5EED0005

EMAIL_PATTERN = r"(?P<user>\w+)@(?P<domain>\w+)\.(?P<tld>\w+)"          # tests/integration/replace_test.go:18
URL_PATTERN = r"(?P<protocol>https?)://(?P<host>[\w\.-]+)(?::(?P<port>\d+))?(?P<path>/[\w\./]*)?"  # scripts/curated/cases.go:67
DATE_PATTERN = r"\d{4}-\d{2}-\d{2}"                                      # BASELINE.json configs[0]
DATE_CAPTURE_PATTERN = r"(\d{4}-\d{2}-\d{2})"                           # tests/integration/streaming/testdata/generate.go:6


def _rng(seed, block):
    return np.random.Generator(np.random.Philox(key=seed, counter=[block, 0, 0, 0]))


class _Pool:
    """A fixed table of byte tokens (rows padded to the longest token)."""

    def __init__(self, tokens):
        self.n = len(tokens)
        self.lens = np.array([len(t) for t in tokens], dtype=np.int32)
        w = int(self.lens.max())
        self.mat = np.zeros((self.n, w), dtype=np.uint8)
        for i, t in enumerate(tokens):
            self.mat[i, :len(t)] = np.frombuffer(t, dtype=np.uint8)
        self.cols = np.arange(w, dtype=np.int32)

    def concat(self, idx, device=None):
        """Concatenate the tokens idx[0], idx[1], ... into one byte array.

        device=None: numpy on the host.  Otherwise a torch device: the gather runs there (same bytes),
        which is what bench.py uses to fill multi-GiB buffers in seconds."""
        if device is None:
            lens = self.lens[idx]
            starts = np.cumsum(lens) - lens
            tok = np.repeat(np.arange(idx.size, dtype=np.int64), lens)
            col = np.arange(tok.size, dtype=np.int64) - starts[tok]
            return self.mat[idx[tok], col]
        import torch
        if not hasattr(self, "_t") or self._t[0].device != torch.device(device):
            self._t = (torch.from_numpy(self.mat).to(device), torch.from_numpy(self.lens.astype(np.int64)).to(device))
        mat, lens_t = self._t
        idx_t = torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int64)).to(device)
        lens = lens_t[idx_t]
        starts = torch.cumsum(lens, 0) - lens
        tok = torch.repeat_interleave(torch.arange(idx_t.numel(), device=device), lens)
        col = torch.arange(tok.numel(), device=device) - starts[tok]
        return mat[idx_t[tok], col]


def _letters(rng, n, lo, hi, alphabet=b"abcdefghijklmnopqrstuvwxyz"):
    a = np.frombuffer(alphabet, dtype=np.uint8)
    out = []
    for _ in range(n):
        k = int(rng.integers(lo, hi + 1))
        out.append(a[rng.integers(0, len(a), size=k)].tobytes())
    return out


_pools = {}


def _url_pool():
    if "url" not in _pools:
        rng = _rng(SEED_C3, 1 << 40)
        words = [w + b" " for w in _letters(rng, 8192, 2, 10)]
        words += [w + b"\n" for w in _letters(rng, 512, 2, 10)]
        urls = []
        for _ in range(4096):
            host = b".".join(_letters(rng, int(rng.integers(1, 3)), 3, 10)) + b"." + _letters(rng, 1, 2, 3)[0]
            u = (b"https" if rng.integers(0, 2) else b"http") + b"://" + host
            if rng.integers(0, 3) == 0:
                u += b":" + str(int(rng.integers(80, 9999))).encode()
            if rng.integers(0, 3) != 0:
                u += b"/" + b"/".join(_letters(rng, int(rng.integers(0, 4)), 1, 8))
            urls.append(u + b" ")
        _pools["url"] = (_Pool(words + urls), len(words), len(urls))
    return _pools["url"]


def url_text_block(block, seed=SEED_C3, url_every=40, device=None):
    """C3: lowercase prose with one URL per ~256 bytes (one token in `url_every` is a URL)."""
    pool, n_words, n_urls = _url_pool()
    rng = _rng(seed, block)
    n_tok = BLOCK // 5 + 64
    idx = rng.integers(0, n_words, size=n_tok)
    is_url = rng.integers(0, url_every, size=n_tok) == 0
    idx[is_url] = n_words + rng.integers(0, n_urls, size=int(is_url.sum()))
    out = pool.concat(idx, device)
    assert out.shape[0] >= BLOCK
    return out[:BLOCK]


def _log_pool():
    if "log" not in _pools:
        rng = _rng(SEED_C2, 1 << 40)
        ts = []
        for _ in range(1024):
            ts.append(("2025-%02d-%02dT%02d:%02d:%02d.%03dZ " % (rng.integers(1, 13), rng.integers(1, 29), rng.integers(0, 24),
                                                               rng.integers(0, 60), rng.integers(0, 60), rng.integers(0, 1000))).encode())
        levels = [b"INFO ", b"WARN ", b"ERROR ", b"DEBUG "]
        words = [w + b" " for w in _letters(rng, 8192, 2, 10)]
        w_alpha = b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789_"
        emails = []
        for _ in range(4096):
            u, d = _letters(rng, 2, 3, 12, w_alpha)
            emails.append(u + b"@" + d + b"." + _letters(rng, 1, 2, 3)[0] + b" ")
        adv = [b"a@b ", b"@x.y ", b"a@b. ", b"a..b@c.d ", b"user@@host.com ", b"x@y.z@w.v ", b"name@host ", b"@ ", b"a@.b "]
        nl = [b"\n"]
        toks = ts + levels + words + emails + adv + nl
        base = {}
        pos = 0
        for name, lst in (("ts", ts), ("level", levels), ("word", words), ("email", emails), ("adv", adv), ("nl", nl)):
            base[name] = (pos, len(lst))
            pos += len(lst)
        _pools["log"] = (_Pool(toks), base)
    return _pools["log"]


def log_text_block(block, seed=SEED_C2, device=None):
    """C2: '<ISO ts> <LEVEL> <5-15 words>[ <email>]\\n' lines (~120 B); email on 25 % of the lines,
    adversarial near-miss fragments on 1 %."""
    pool, base = _log_pool()
    rng = _rng(seed, block)
    n_lines = BLOCK // 60 + 16
    slots = np.full((n_lines, 20), -1, dtype=np.int64)
    slots[:, 0] = base["ts"][0] + rng.integers(0, base["ts"][1], size=n_lines)
    slots[:, 1] = base["level"][0] + rng.integers(0, base["level"][1], size=n_lines)
    n_words = rng.integers(5, 16, size=n_lines)
    w = base["word"][0] + rng.integers(0, base["word"][1], size=(n_lines, 15))
    w[np.arange(15)[None, :] >= n_words[:, None]] = -1
    slots[:, 2:17] = w
    r = rng.integers(0, 100, size=n_lines)
    em = base["email"][0] + rng.integers(0, base["email"][1], size=n_lines)
    slots[:, 17] = np.where(r < 25, em, -1)
    ad = base["adv"][0] + rng.integers(0, base["adv"][1], size=n_lines)
    slots[:, 18] = np.where(r == 99, ad, -1)
    slots[:, 19] = base["nl"][0]
    idx = slots.reshape(-1)
    out = pool.concat(idx[idx >= 0], device)
    assert out.shape[0] >= BLOCK
    return out[:BLOCK]


def patterned_stream_block(block, seed=SEED_C5, every=50, digit_noise=0.0, device=None):
    """C5: PatternedReader analogue (tests/integration/streaming/memory_test.go:23-48): the literal
    '2024-01-15' every `every` bytes over noise from "abcdefghijk \\n\\t"; `digit_noise` replaces that
    fraction of the noise bytes with digits / dashes (fires the skip-restart quirk)."""
    rng = _rng(seed, block)
    alpha = np.frombuffer(b"abcdefghijk \n\t", dtype=np.uint8)
    out = alpha[rng.integers(0, alpha.size, size=BLOCK)]
    if digit_noise > 0:
        dn = np.frombuffer(b"0123456789-", dtype=np.uint8)
        m = rng.random(BLOCK) < digit_noise
        out[m] = dn[rng.integers(0, dn.size, size=int(m.sum()))]
    date = np.frombuffer(b"2024-01-15", dtype=np.uint8)
    # absolute positions k*every (k >= 1) that start inside this block; a date may spill into the
    # next block, which re-creates its tail from the same rule
    g0 = block * BLOCK
    first = ((g0 - 9 + every - 1) // every) * every if g0 > 0 else every
    p = np.arange(max(first, every), g0 + BLOCK, every, dtype=np.int64) - g0
    for j in range(10):
        q = p + j
        ok = (q >= 0) & (q < BLOCK)
        out[q[ok]] = date[j]
    if device is not None:
        import torch
        return torch.from_numpy(out).to(device)
    return out


def make_buffer(kind, n_bytes, first_block=0, device=None, **kw):
    """Blocks first_block.. concatenated to exactly n_bytes: a host numpy array, or (device given) a
    torch uint8 tensor filled block by block on that device."""
    fn = {"url": url_text_block, "log": log_text_block, "stream": patterned_stream_block}[kind]
    if device is None:
        out = np.empty(n_bytes, dtype=np.uint8)
    else:
        import torch
        out = torch.empty(n_bytes, dtype=torch.uint8, device=device)
    for b in range((n_bytes + BLOCK - 1) // BLOCK):
        lo = b * BLOCK
        hi = min(lo + BLOCK, n_bytes)
        out[lo:hi] = fn(first_block + b, device=device, **kw)[:hi - lo]
    return out


def date_strings(n=10000, seed=SEED_C1):
    """C1: N ASCII strings, length 0..64: 35 % contain a valid date in lowercase noise, 20 % are
    'near-miss then date' (1-7 digits glued in front of a date: fires SURVEY Q1), 15 % digits/dashes
    only, 20 % random printable ASCII, 10 % empty or shorter than 10 bytes."""
    rng = _rng(seed, 0)
    out = []
    low = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz ", dtype=np.uint8)
    dd = np.frombuffer(b"0123456789-", dtype=np.uint8)
    for _ in range(n):
        r = int(rng.integers(0, 100))
        date = ("%04d-%02d-%02d" % (rng.integers(1900, 2100), rng.integers(1, 13), rng.integers(1, 29))).encode()
        if r < 35:
            a = low[rng.integers(0, low.size, size=int(rng.integers(0, 28)))].tobytes()
            b = low[rng.integers(0, low.size, size=int(rng.integers(0, 27)))].tobytes()
            s = a + date + b
        elif r < 55:
            a = low[rng.integers(0, low.size, size=int(rng.integers(0, 20)))].tobytes()
            k = int(rng.integers(1, 8))
            pre = dd[rng.integers(0, 10, size=k)].tobytes()
            if rng.integers(0, 3) == 0:
                pre = date[:int(rng.integers(1, 8))]
            s = a + pre + date + low[rng.integers(0, low.size, size=int(rng.integers(0, 10)))].tobytes()
        elif r < 70:
            s = dd[rng.integers(0, dd.size, size=int(rng.integers(0, 65)))].tobytes()
        elif r < 90:
            s = rng.integers(32, 127, size=int(rng.integers(0, 65))).astype(np.uint8).tobytes()
        else:
            s = rng.integers(32, 127, size=int(rng.integers(0, 10))).astype(np.uint8).tobytes()
        out.append(s[:64])
    return out


def mutate_inputs(inputs, n, seed=SEED_C4, stream=0):
    """C4: derive n short ASCII inputs from a pattern's corpus inputs: keep 40 %, substitute one or two
    bytes 30 %, truncate / extend 30 %."""
    rng = _rng(seed, stream)
    base = [x.encode("utf-8") if isinstance(x, str) else bytes(x) for x in inputs]
    base = [b for b in base if all(c < 128 for c in b)] or [b""]
    out = []
    for _ in range(n):
        s = bytearray(base[int(rng.integers(0, len(base)))])
        r = int(rng.integers(0, 100))
        if r >= 40 and r < 70 and len(s):
            for _k in range(int(rng.integers(1, 3))):
                s[int(rng.integers(0, len(s)))] = int(rng.integers(32, 127))
        elif r >= 70:
            if rng.integers(0, 2) and len(s):
                s = s[:int(rng.integers(0, len(s) + 1))]
            else:
                s += bytes(rng.integers(32, 127, size=int(rng.integers(1, 9))).astype(np.uint8))
        out.append(bytes(s))
    return out
