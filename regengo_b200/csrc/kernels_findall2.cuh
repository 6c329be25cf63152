// Shared definitions of the segment-slab FindAll kernels (kernels_scan6.cuh, kernels_btrun.cuh,
// kernels_chain.cuh, kernels_emit.cuh): segment size, key encoding.
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

}  // namespace rgx
