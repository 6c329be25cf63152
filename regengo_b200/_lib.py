"""ctypes binding of libregengo_b200.so (the C ABI in include/regengo_b200.h).

The library is built in-tree by regengo_b200/build.py.  Importing this module never falls back to
a Python matcher: if the shared library is missing it is built, and if that fails the import fails.
"""
import ctypes as C
import os

from . import build as _build

RGX_OK = 0
RGX_EINVAL, RGX_EPATTERN, RGX_EUNSUPPORTED, RGX_ECUDA, RGX_ENOMEM, RGX_EBUFFER_TOO_SMALL, RGX_ECAPACITY = -1, -2, -3, -4, -5, -6, -7


class Options(C.Structure):
    _fields_ = [("force_thompson", C.c_int32), ("force_tnfa", C.c_int32), ("force_tdfa", C.c_int32), ("tdfa_threshold", C.c_int32)]


class Info(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_inst", "num_cap", "n_groups", "match_engine", "find_engine", "match_memo", "find_memo", "per_capture_ckpt",
        "anchored", "min_match_len", "max_match_len", "default_max_leftover", "min_buffer", "tdfa_states", "tdfa_tags")] + [
        ("reserved", C.c_int32 * 5)]


_P = C.c_void_p
_SIGS = {
    "rgx_compile": (C.c_int, [C.c_char_p, C.POINTER(Options), C.POINTER(_P)]),
    "rgx_program_free": (None, [_P]),
    "rgx_program_info": (C.c_int, [_P, C.POINTER(Info)]),
    "rgx_program_json": (C.c_char_p, [_P]),
    "rgx_program_device_plan": (C.c_int64, [_P, C.c_char_p, C.c_size_t]),
    "rgx_program_device_image": (C.c_int64, [_P, _P, C.c_size_t]),
    "rgx_program_group_name": (C.c_char_p, [_P, C.c_int32]),
    "rgx_program_blob": (C.c_int64, [_P, _P, C.c_size_t]),
    "rgx_load": (C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    "rgx_last_error": (C.c_char_p, []),
    "rgx_version": (C.c_char_p, []),
    "rgx_ctx_create": (C.c_int, [C.c_int32, C.POINTER(_P)]),
    "rgx_ctx_destroy": (None, [_P]),
    "rgx_ctx_launches": (C.c_int64, [_P]),
    "rgx_ctx_stat": (C.c_int64, [_P, C.c_int32]),
    "rgx_ctx_set_chunk_bytes": (C.c_int, [_P, C.c_uint64]),
    "rgx_ctx_enable_timing": (C.c_int, [_P, C.c_int32]),
    "rgx_ctx_last_timing": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "rgx_ctx_stream": (_P, [_P]),
    "rgx_ctx_sync": (C.c_int, [_P]),
    "rgx_match_batch": (C.c_int, [_P, _P, _P, _P, C.c_uint64, _P]),
    "rgx_match_batch_dev": (C.c_int, [_P, _P, _P, _P, C.c_uint64, _P]),
    "rgx_match_multi": (C.c_int, [_P, _P, C.c_uint32, _P, _P, _P, _P]),
    "rgx_match_multi_dev": (C.c_int, [_P, _P, C.c_uint32, _P, _P, _P, _P]),
    "rgx_find_batch": (C.c_int, [_P, _P, _P, _P, C.c_uint64, _P, _P]),
    "rgx_find_batch_dev": (C.c_int, [_P, _P, _P, _P, C.c_uint64, _P, _P]),
    "rgx_replace_batch": (C.c_int, [_P, _P, C.c_char_p, C.c_uint64, _P, _P, C.c_uint64, _P, C.c_uint64, _P, C.POINTER(C.c_uint64)]),
    "rgx_replace_batch_dev": (C.c_int, [_P, _P, C.c_char_p, C.c_uint64, _P, _P, C.c_uint64, _P, C.c_uint64, _P, C.POINTER(C.c_uint64)]),
    "rgx_replace_template_check": (C.c_int, [_P, C.c_char_p, C.c_uint64]),
    "rgx_replace_template_dump": (C.c_int64, [_P, C.c_char_p, C.c_uint64, C.c_char_p, C.c_size_t]),
    "rgx_find_all": (C.c_int64, [_P, _P, _P, C.c_uint64, C.c_int64, _P, C.c_uint64]),
    "rgx_find_all_rle": (C.c_int64, [_P, _P, _P, C.c_uint64, C.c_int64, _P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "rgx_find_all_dev": (C.c_int64, [_P, _P, _P, C.c_uint64, C.c_int64, _P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "rgx_find_all_shard_dev": (C.c_int64, [_P, _P, _P, C.c_uint64, C.c_uint64, C.c_int32, C.c_int64, C.c_int64, C.c_int32, _P, _P,
                                           C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]),
    "rgx_find_all_shard_pre_dev": (C.c_int64, [_P, _P, _P, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32, C.c_int64, _P, _P,
                                               C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "rgx_find_reader": (C.c_int64, [_P, _P, _P, C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _P, _P, _P, C.c_uint64]),
    "rgx_find_reader_dev": (C.c_int64, [_P, _P, _P, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                        _P, _P, _P, C.c_uint64]),
    "rgx_stream_config": (C.c_int, [_P, C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "rgx_dev_alloc": (C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    "rgx_dev_free": (C.c_int, [_P, _P]),
    "rgx_dev_upload": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "rgx_dev_download": (C.c_int, [_P, _P, _P, C.c_size_t]),
}

_lib = None


def lib_path():
    return _build.LIB


def load():
    """Load (building first if needed) the shared library and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_build.LIB):
        _build.build()
    lib = C.CDLL(_build.LIB)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError here == the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIGS)


class RegengoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


class BufferTooSmall(RegengoError):
    """stream.ErrBufferTooSmall (stream/stream.go:80-93)."""


def check(rc):
    if rc is not None and rc < 0:
        msg = load().rgx_last_error().decode("utf-8", "replace")
        if rc == RGX_EBUFFER_TOO_SMALL:
            raise BufferTooSmall(rc, msg)
        raise RegengoError(rc, msg)
    return rc
