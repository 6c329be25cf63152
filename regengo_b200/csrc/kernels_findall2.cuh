// Shared definitions of the segment-slab FindAll kernels (kernels_scan5.cuh, kernels_btrun.cuh,
// kernels_chain.cuh, kernels_emit.cuh): segment size, key encoding, SWAR byte compare.
//
// Slab entry j of a segment is CANDIDATE j (key.y == KEY_INVALID when it did not match), so no
// in-order compaction is needed in the scan; the chain/emit kernels skip invalid entries.
#pragma once
#include "engines.cuh"
#include "kernels_findall.cuh"

namespace rgx {

constexpr uint32_t SEG2_BYTES = 32768;    // bytes per segment (one warp) of the fast TDFA scan
constexpr uint32_t KEY_INVALID = 0xFFFFFFFFu;
constexpr int NT_MAX2 = 16;               // tags handled by the fast walk

// approximate per-byte equality: bit 7 of every byte of x that equals pat's byte is set; a byte
// above a matching byte may be flagged too (false positives only -- candidates are verified)
__device__ __forceinline__ uint32_t eq_approx(uint32_t w, uint32_t pat) {
  const uint32_t x = w ^ pat;
  return (x - 0x01010101u) & ~x & 0x80808080u;
}

}  // namespace rgx
