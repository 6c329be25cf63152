"""regengo_b200 -- B200-native matching path for regengo patterns (MatchBytes / FindBytes /
FindAllBytes / FindReader) behind a C ABI.  See DESIGN.md and include/regengo_b200.h."""
from .api import Pattern, Result, StreamConfig, StreamMatch, compile, context, launches, match_multi, pack_inputs  # noqa: F401
from ._lib import BufferTooSmall, RegengoError  # noqa: F401

__all__ = ["Pattern", "Result", "StreamConfig", "StreamMatch", "compile", "context", "launches", "match_multi", "pack_inputs",
           "RegengoError", "BufferTooSmall"]
