"""Host-side mirror of the method set regengo generates for one pattern.

`regengo.Compile(Options{Pattern, Name, ...})` (regengo.go:86-156) emits a Go type with
MatchBytes / FindBytes / FindAllBytes / FindReader / MatchLengthInfo / DefaultMaxLeftover
(internal/compiler/compiler.go:204-367).  `Pattern` below exposes the same methods (snake_case) with
the same argument meaning and error behaviour; every matching call goes through the C ABI
(include/regengo_b200.h) to the CUDA kernels.  There is no Python or CPU matcher behind it.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import BufferTooSmall, RegengoError, check  # noqa: F401

_contexts = {}


def context(device=0):
    """One device context (CUDA stream + scratch) per device ordinal, created on first use."""
    if device not in _contexts:
        L = _lib.load()
        h = C.c_void_p()
        check(L.rgx_ctx_create(device, C.byref(h)))
        _contexts[device] = h
    return _contexts[device]


def launches(device=0):
    return int(_lib.load().rgx_ctx_launches(context(device)))


@dataclass
class StreamConfig:
    """stream.Config (stream/stream.go:21-39)."""
    buffer_size: int = 0
    max_leftover: int = 0


@dataclass
class StreamMatch:
    """stream.Match[T] (stream/stream.go:66-79)."""
    result: "Result"
    stream_offset: int
    chunk_index: int


class Result:
    """<Name>BytesResult (captures.go:83-118): zero-copy views into the caller's input."""
    __slots__ = ("_data", "_rec", "_names")

    def __init__(self, data, rec, names):
        self._data, self._rec, self._names = data, rec, names

    @property
    def offsets(self):
        return [int(x) for x in self._rec]

    def span(self, g=0):
        return int(self._rec[2 * g]), int(self._rec[2 * g + 1])

    def group(self, g=0):
        """bytes of group g, or None where the reference leaves the field nil."""
        if isinstance(g, str):
            g = self._names.index(g)
        s, e = self.span(g)
        if s < 0:
            return None
        return bytes(self._data[s:e])

    @property
    def match(self):
        return self.group(0)

    def groups(self):
        return [self.group(g) for g in range(1, len(self._rec) // 2)]

    def __repr__(self):
        return f"Result(match={self.match!r}, groups={self.groups()!r})"


def _as_u8(data):
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(bytes(data) if not isinstance(data, (bytes, bytearray, memoryview)) else data, dtype=np.uint8)


def pack_inputs(inputs):
    """list of bytes -> (bytes[], offsets[n+1] uint64), the batch layout of the C ABI."""
    lens = np.fromiter((len(x) for x in inputs), dtype=np.uint64, count=len(inputs))
    offs = np.zeros(len(inputs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offs[1:])
    data = np.frombuffer(b"".join(bytes(x) for x in inputs), dtype=np.uint8) if len(inputs) else np.zeros(0, np.uint8)
    return data, offs


def match_multi(patterns, data, offsets, prog_first, device=0):
    """Batched MatchBytes of MANY patterns in one launch (rgx_match_multi): inputs prog_first[p] .. prog_first[p+1]-1 of
    the packed batch (uint8 data, uint64 offsets[n+1]) are matched against patterns[p].  -> uint8[n]"""
    data = _as_u8(data)
    offs = np.ascontiguousarray(offsets, dtype=np.uint64)
    pf = np.ascontiguousarray(prog_first, dtype=np.uint64)
    assert pf.size == len(patterns) + 1
    n = int(pf[-1] - pf[0])
    out = np.zeros(n, dtype=np.uint8)
    handles = (C.c_void_p * len(patterns))(*[p._h for p in patterns])
    if n:
        check(_lib.load().rgx_match_multi(context(device), handles, len(patterns), data.ctypes.data, offs.ctypes.data,
                                          pf.ctypes.data, out.ctypes.data))
    return out


class Pattern:
    """A compiled pattern == one generated regengo type."""

    def __init__(self, pattern, name=None, force_thompson=False, force_tnfa=False, force_tdfa=False, tdfa_threshold=0,
                 device=0, _blob=None):
        L = _lib.load()
        self._h = C.c_void_p()
        self.pattern = pattern
        self.name = name or "Pattern"
        self.device = device
        if _blob is not None:
            buf = bytes(_blob)
            check(L.rgx_load(buf, len(buf), C.byref(self._h)))
        else:
            o = _lib.Options(int(force_thompson), int(force_tnfa), int(force_tdfa), int(tdfa_threshold))
            pat = pattern.encode("utf-8") if isinstance(pattern, str) else bytes(pattern)
            check(L.rgx_compile(pat, C.byref(o), C.byref(self._h)))
        info = _lib.Info()
        check(L.rgx_program_info(self._h, C.byref(info)))
        self.info = info
        self.num_cap = info.num_cap
        self.group_names = [""] + [(L.rgx_program_group_name(self._h, i) or b"").decode() for i in range(1, info.n_groups + 1)]

    @classmethod
    def from_blob(cls, blob, device=0):
        return cls(None, device=device, _blob=blob)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.load().rgx_program_free(h)
            except Exception:
                pass
            self._h = None

    # ---- inspection (compiler.go:705-713, streaming.go:41-46) ----
    def blob(self):
        L = _lib.load()
        n = L.rgx_program_blob(self._h, None, 0)
        buf = C.create_string_buffer(n)
        L.rgx_program_blob(self._h, buf, n)
        return buf.raw

    def json(self):
        return _lib.load().rgx_program_json(self._h).decode("utf-8")

    def match_length_info(self):
        return self.info.min_match_len, self.info.max_match_len

    def default_max_leftover(self):
        return self.info.default_max_leftover

    def stream_config(self, cfg=None):
        cfg = cfg or StreamConfig()
        b, l = C.c_int64(), C.c_int64()
        check(_lib.load().rgx_stream_config(self._h, cfg.buffer_size, cfg.max_leftover, C.byref(b), C.byref(l)))
        return StreamConfig(b.value, l.value)

    # ---- MatchBytes / MatchString ----
    def match_bytes(self, data):
        return bool(self.match_batch([data])[0])

    def match_string(self, s):
        return self.match_bytes(s.encode("utf-8"))

    def match_batch(self, inputs, offsets=None):
        """Batched MatchBytes: list of bytes, or (uint8 array, uint64 offsets[n+1]).  -> uint8[n]"""
        data, offs = (pack_inputs(inputs) if offsets is None else (_as_u8(inputs), np.ascontiguousarray(offsets, dtype=np.uint64)))
        n = offs.size - 1
        out = np.zeros(n, dtype=np.uint8)
        if n:
            check(_lib.load().rgx_match_batch(context(self.device), self._h, data.ctypes.data, offs.ctypes.data, n, out.ctypes.data))
        return out

    # ---- FindBytes ----
    def find_batch(self, inputs, offsets=None):
        """Batched FindBytes -> (found uint8[n], records int64[n, num_cap] relative to each input)."""
        data, offs = (pack_inputs(inputs) if offsets is None else (_as_u8(inputs), np.ascontiguousarray(offsets, dtype=np.uint64)))
        n = offs.size - 1
        found = np.zeros(n, dtype=np.uint8)
        rec = np.full((n, self.num_cap), -1, dtype=np.int64)
        if self.info.find_engine == 0 or n:
            check(_lib.load().rgx_find_batch(context(self.device), self._h, data.ctypes.data, offs.ctypes.data, n,
                                             found.ctypes.data, rec.ctypes.data))
        return found, rec

    def find_bytes(self, data):
        """(*Result, ok) of FindBytes -> Result or None."""
        data = bytes(data)
        found, rec = self.find_batch([data])
        return Result(data, rec[0], self.group_names) if found[0] else None

    # ---- ReplaceAllBytesAppend ----
    def replace_all_batch(self, inputs, template, offsets=None):
        """Batched ReplaceAllBytesAppend (replace.go:192-273) -> (out_bytes uint8[], out_offs uint64[n + 1])."""
        data, offs = (pack_inputs(inputs) if offsets is None else (_as_u8(inputs), np.ascontiguousarray(offsets, dtype=np.uint64)))
        t = template.encode("utf-8") if isinstance(template, str) else bytes(template)
        n = offs.size - 1
        L = _lib.load()
        out_offs = np.zeros(n + 1, dtype=np.uint64)
        total = C.c_uint64()
        cap = int(data.size) + 64 * max(n, 1)
        for _ in range(2):
            out = np.empty(max(cap, 1), dtype=np.uint8)
            rc = L.rgx_replace_batch(context(self.device), self._h, t, len(t), data.ctypes.data, offs.ctypes.data, n, out.ctypes.data,
                                     cap, out_offs.ctypes.data, C.byref(total))
            if rc != _lib.RGX_ECAPACITY:
                break
            cap = int(total.value)
        check(rc)
        return out[: int(total.value)], out_offs

    def replace_all(self, data, template):
        """ReplaceAllBytes(input, template) -> bytes."""
        out, _ = self.replace_all_batch([bytes(data)], template)
        return out.tobytes()

    def template_segments(self, template):
        """replace.Parse + name/index resolution (host only): [(group or -1, literal bytes)]."""
        import json as _json
        t = template.encode("utf-8") if isinstance(template, str) else bytes(template)
        L = _lib.load()
        n = check(L.rgx_replace_template_dump(self._h, t, len(t), None, 0))
        buf = C.create_string_buffer(int(n) + 1)
        check(L.rgx_replace_template_dump(self._h, t, len(t), buf, int(n) + 1))
        return [(g, bytes.fromhex(h)) for g, h in _json.loads(buf.value.decode())]

    def device_plan(self):
        """How the device kernels will run this pattern (host computation; dict)."""
        import json as _json
        L = _lib.load()
        n = check(L.rgx_program_device_plan(self._h, None, 0))
        buf = C.create_string_buffer(int(n) + 1)
        check(L.rgx_program_device_plan(self._h, buf, int(n) + 1))
        return _json.loads(buf.value.decode())

    def device_image(self):
        """The packed device image (numpy uint32); device_plan()'s w6_* entries are offsets into it."""
        L = _lib.load()
        n = int(check(L.rgx_program_device_image(self._h, None, 0)))
        words = np.empty(n, dtype=np.uint32)
        check(L.rgx_program_device_image(self._h, words.ctypes.data, n))
        return words

    # ---- FindAllBytes ----
    def find_all_offsets(self, data, n=-1):
        """FindAllBytes(data, n) as (count, int64[count, num_cap])."""
        a = _as_u8(data)
        L = _lib.load()
        cap = max(64, a.size // 32 + 64)
        for _ in range(2):
            out = np.empty((cap, self.num_cap), dtype=np.int64)
            cnt = check(L.rgx_find_all(context(self.device), self._h, a.ctypes.data, a.size, n, out.ctypes.data, cap))
            if cnt <= cap:
                return int(cnt), out[:cnt]
            cap = int(cnt)
        raise RegengoError(_lib.RGX_ECAPACITY, "FindAll output did not fit")

    def find_all_bytes(self, data, n=-1):
        data = bytes(data)
        _, recs = self.find_all_offsets(data, n)
        return [Result(data, r, self.group_names) for r in recs]

    # ---- FindReader ----
    def find_reader_offsets(self, data, cfg=None, first_chunk=0, n_chunks=-1):
        """FindReader over bytes.Reader(data) -> (count, stream_off[], chunk_idx[], records[])."""
        cfg = cfg or StreamConfig()
        a = _as_u8(data)
        L = _lib.load()
        cap = max(64, a.size // 32 + 64)
        for _ in range(2):
            so = np.empty(cap, dtype=np.int64)
            ci = np.empty(cap, dtype=np.int32)
            out = np.empty((cap, self.num_cap), dtype=np.int64)
            cnt = check(L.rgx_find_reader(context(self.device), self._h, a.ctypes.data, a.size, cfg.buffer_size, cfg.max_leftover,
                                          first_chunk, n_chunks, so.ctypes.data, ci.ctypes.data, out.ctypes.data, cap))
            if cnt <= cap:
                return int(cnt), so[:cnt], ci[:cnt], out[:cnt]
            cap = int(cnt)
        raise RegengoError(_lib.RGX_ECAPACITY, "FindReader output did not fit")

    def find_reader(self, reader, cfg, on_match):
        """FindReader(r io.Reader, cfg stream.Config, onMatch func(stream.Match) bool) error.

        `reader` is drained with .read() (the device path needs the stream resident); matches are
        delivered in order and delivery stops when on_match returns False (streaming.go:209-217)."""
        data = reader.read() if hasattr(reader, "read") else bytes(reader)
        if isinstance(data, str):
            data = data.encode("utf-8")
        n, so, ci, recs = self.find_reader_offsets(data, cfg)
        for i in range(n):
            if not on_match(StreamMatch(Result(data, recs[i], self.group_names), int(so[i]), int(ci[i]))):
                break
        return None

    def find_reader_count(self, reader, cfg=None):
        """FindReaderCount(r, cfg) (int64, error): the number of matches (streaming.go:258-283)."""
        data = reader.read() if hasattr(reader, "read") else bytes(reader)
        return self.find_reader_offsets(data, cfg)[0]

    def find_reader_first(self, reader, cfg=None):
        """FindReaderFirst(r, cfg) (*TBytesResult, int64, error): the first match and its StreamOffset, or (None, 0)
        (streaming.go:285-313: FindReader with a callback that copies the result and returns false)."""
        data = reader.read() if hasattr(reader, "read") else bytes(reader)
        if isinstance(data, str):
            data = data.encode("utf-8")
        n, so, _, recs = self.find_reader_offsets(data, cfg)
        if n == 0:
            return None, 0
        return Result(data, recs[0], self.group_names), int(so[0])


def compile(pattern, **kw):  # noqa: A001  (mirrors regengo.Compile)
    return Pattern(pattern, **kw)
