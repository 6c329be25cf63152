"""Shared helpers for the test-suite."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from regengo_b200 import _lib  # noqa: E402


def compile_json(pattern, **opts):
    """rgx_compile -> parsed rgx_program_json (front-end only, no GPU)."""
    L = _lib.load()
    p = C.c_void_p()
    o = _lib.Options(**{k: int(v) for k, v in opts.items()})
    rc = L.rgx_compile(pattern.encode("utf-8") if isinstance(pattern, str) else pattern, C.byref(o), C.byref(p))
    if rc != 0:
        raise ValueError(L.rgx_last_error().decode())
    try:
        return json.loads(L.rgx_program_json(p).decode("utf-8"))
    finally:
        L.rgx_program_free(p)


def compile_blob(pattern, **opts):
    L = _lib.load()
    p = C.c_void_p()
    o = _lib.Options(**{k: int(v) for k, v in opts.items()})
    rc = L.rgx_compile(pattern.encode("utf-8") if isinstance(pattern, str) else pattern, C.byref(o), C.byref(p))
    if rc != 0:
        raise ValueError(L.rgx_last_error().decode())
    try:
        n = L.rgx_program_blob(p, None, 0)
        buf = C.create_string_buffer(n)
        L.rgx_program_blob(p, buf, n)
        return buf.raw
    finally:
        L.rgx_program_free(p)
