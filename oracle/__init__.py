"""CPU oracle for the regengo matching path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  It wraps oracle/rgx_oracle.c (a plain-C restatement of the loops regengo generates,
each function citing the reference file:line it follows) through ctypes.  The product package
regengo_b200 never imports it.

The oracle consumes a program blob (regengo_b200/csrc/blob.hpp).  Blobs come either from the
product's front-end (rgx_compile -> rgx_program_blob) or -- for the strongest pin -- straight from
the programs mined out of the reference's own generated Go files (tests/golden/golden_blob.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "rgx_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "-B", "liboracle.so"])
    return LIB


def _load():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        P = C.c_void_p
        L.orc_open.restype = P
        L.orc_open.argtypes = [P, C.c_size_t]
        L.orc_close.argtypes = [P]
        L.orc_num_cap.argtypes = [P]
        L.orc_match.argtypes = [P, P, C.c_int64]
        L.orc_find.argtypes = [P, P, C.c_int64, P]
        L.orc_find_all.restype = C.c_int64
        L.orc_find_all.argtypes = [P, P, C.c_int64, C.c_int64, P, C.c_int64]
        L.orc_stream_config.argtypes = [P, C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_find_reader.restype = C.c_int64
        L.orc_find_reader.argtypes = [P, P, C.c_int64, C.c_int64, C.c_int64, P, P, P, C.c_int64]
        L.orc_match_batch.argtypes = [P, P, P, C.c_uint64, P]
        L.orc_find_batch.argtypes = [P, P, P, C.c_uint64, P, P]
        L.orc_replace_all.restype = C.c_int64
        L.orc_replace_all.argtypes = [P, P, C.c_int64, P, C.c_int64, P, C.c_int64]
        L.orc_replace_batch.restype = C.c_int64
        L.orc_replace_batch.argtypes = [P, P, P, C.c_uint64, P, C.c_int64, P, C.c_int64, P]
        _lib = L
    return _lib


def _as_u8(data):
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(bytes(data), dtype=np.uint8)


class Oracle:
    """One pattern, given as a program blob (bytes / uint32 array)."""

    def __init__(self, blob):
        L = _load()
        self._blob = np.frombuffer(bytes(blob), dtype=np.uint32).copy()
        self._h = L.orc_open(self._blob.ctypes.data, self._blob.size)
        if not self._h:
            raise ValueError("bad program blob")
        self.num_cap = L.orc_num_cap(self._h)
        self.find_engine = int(self._blob[9])
        self.match_engine = int(self._blob[8])

    def __del__(self):
        if getattr(self, "_h", None):
            _load().orc_close(self._h)
            self._h = None

    def match(self, data):
        a = _as_u8(data)
        return bool(_load().orc_match(self._h, a.ctypes.data, a.size))

    def find(self, data):
        """FindBytes: None or the offset record (list of 2*(k+1) ints, -1 pairs for nil groups)."""
        a = _as_u8(data)
        out = np.empty(self.num_cap, dtype=np.int64)
        r = _load().orc_find(self._h, a.ctypes.data, a.size, out.ctypes.data)
        if r < 0:
            raise RuntimeError("Find* is not generated for a pattern without captures")
        return out.tolist() if r == 1 else None

    def find_all(self, data, n=-1, cap=None, count_only=False):
        """FindAllBytes(data, n): (count, int64 array [min(count,cap), num_cap]).  count_only: one pass, the count and
        whatever fitted `cap` (timing runs)."""
        a = _as_u8(data)
        L = _load()
        if cap is None:
            cap = max(16, a.size // 4 + 16)
        out = np.empty((cap, self.num_cap), dtype=np.int64)
        cnt = L.orc_find_all(self._h, a.ctypes.data, a.size, n, out.ctypes.data, cap)
        if cnt < 0:
            raise RuntimeError("FindAll* is not generated for a pattern without captures")
        if count_only:
            return int(cnt), out[:min(cnt, cap)]
        if cnt > cap:
            return self.find_all(data, n, cap=int(cnt))
        return int(cnt), out[:cnt]

    def stream_config(self, buffer_size=0, max_leftover=0):
        b, l = C.c_int64(), C.c_int64()
        rc = _load().orc_stream_config(self._h, buffer_size, max_leftover, C.byref(b), C.byref(l))
        if rc < 0:
            raise ValueError("stream: buffer size too small")
        return b.value, l.value

    def find_reader(self, data, buffer_size=0, max_leftover=0, cap=None, count_only=False):
        """FindReader over bytes.Reader(data): (count, stream_off[], chunk_idx[], records[])."""
        a = _as_u8(data)
        L = _load()
        if cap is None:
            cap = max(16, a.size // 4 + 16)
        so = np.empty(cap, dtype=np.int64)
        ci = np.empty(cap, dtype=np.int32)
        out = np.empty((cap, self.num_cap), dtype=np.int64)
        cnt = L.orc_find_reader(self._h, a.ctypes.data, a.size, buffer_size, max_leftover, so.ctypes.data, ci.ctypes.data,
                                out.ctypes.data, cap)
        if cnt == -6:
            raise ValueError("stream: buffer size too small")
        if count_only:
            k = min(cnt, cap)
            return int(cnt), so[:k], ci[:k], out[:k]
        if cnt > cap:
            return self.find_reader(data, buffer_size, max_leftover, cap=int(cnt))
        return int(cnt), so[:cnt], ci[:cnt], out[:cnt]

    def match_batch(self, bytes_arr, offs):
        b = _as_u8(bytes_arr)
        o = np.ascontiguousarray(offs, dtype=np.uint64)
        n = o.size - 1
        out = np.empty(n, dtype=np.uint8)
        _load().orc_match_batch(self._h, b.ctypes.data, o.ctypes.data, n, out.ctypes.data)
        return out

    def replace_all(self, data, template):
        """ReplaceAllBytes(input, template) -> bytes (ValueError on a malformed template: the reference panics)."""
        out, _ = self.replace_batch(_as_u8(data), np.array([0, len(data)], dtype=np.uint64), template)
        return out.tobytes()

    def replace_batch(self, bytes_arr, offs, template):
        """ReplaceAllBytesAppend per input -> (out_bytes, out_offs[n + 1])."""
        b = _as_u8(bytes_arr)
        o = np.ascontiguousarray(offs, dtype=np.uint64)
        t = np.frombuffer(template.encode("utf-8") if isinstance(template, str) else bytes(template), dtype=np.uint8)
        n = o.size - 1
        out_offs = np.zeros(n + 1, dtype=np.uint64)
        cap = int(b.size) + 64 * max(n, 1)
        for _ in range(2):
            out = np.empty(max(cap, 1), dtype=np.uint8)
            r = _load().orc_replace_batch(self._h, b.ctypes.data, o.ctypes.data, n, t.ctypes.data if t.size else None, t.size,
                                          out.ctypes.data, cap, out_offs.ctypes.data)
            if r == -7:
                raise ValueError("regengo: invalid replace template")
            if r < 0:
                raise NotImplementedError(f"oracle: template not decided by this restatement ({r})")
            if r <= cap:
                break
            cap = int(r)
        return out[:r], out_offs

    def find_batch(self, bytes_arr, offs):
        b = _as_u8(bytes_arr)
        o = np.ascontiguousarray(offs, dtype=np.uint64)
        n = o.size - 1
        found = np.empty(n, dtype=np.uint8)
        out = np.full((n, self.num_cap), -1, dtype=np.int64)
        _load().orc_find_batch(self._h, b.ctypes.data, o.ctypes.data, n, found.ctypes.data, out.ctypes.data)
        return found, out
