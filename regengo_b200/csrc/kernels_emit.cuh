// findall_emit3_kernel -- ordered output of the kept records: one warp per segment.
// Each group of 32 slab entries is turned into int64 offset records in shared memory (one record per
// lane, at its rank among the kept entries) and then written out as one contiguous, fully coalesced
// run of 8-byte words -- consecutive kept records are consecutive in the output.
// n_limit < 0: everything; otherwise the expanded list is cut after n_limit matches (a record's reps
// are clipped, later records dropped).  *n_written = number of records kept.
#pragma once
#include "kernels_chain.cuh"

namespace rgx {

constexpr int EMIT_WARPS = 8;

template <int ENGINE>
__global__ void __launch_bounds__(EMIT_WARPS * 32) findall_emit3_kernel(
    const DevMeta m, const uint64_t n_seg, const uint32_t seg_bytes, const uint32_t G, const uint32_t mis, const uint64_t len,
    const FindAllBufs fb, const uint32_t* __restrict__ seg_sel, const unsigned long long* __restrict__ seg_reps,
    const unsigned long long* __restrict__ sel_base, const unsigned long long* __restrict__ reps_base,
    const unsigned long long* __restrict__ totals, const long long n_limit, int64_t* __restrict__ out,
    uint32_t* __restrict__ out_reps, const uint64_t cap_records, unsigned long long* n_written, const uint64_t skip_seg,
    const int* err) {
  extern __shared__ __align__(16) long long stage_all[];   // [EMIT_WARPS][32 * nc]
  if (*(const volatile int*)err) return;   // the attempt failed upstream (the host retries): the selections are not valid
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // segments before skip_seg (a whole number of parts: the pre-halo of a shard) report nothing, and what they kept is
  // taken off every output position
  const uint64_t seg = skip_seg + (uint64_t)blockIdx.x * EMIT_WARPS + warp;
  const int nc = ENGINE == FIND_TDFA ? m.t_ntags : m.num_cap;
  long long* stage = stage_all + (size_t)warp * 32 * nc;
  if (seg == skip_seg && lane == 0 && (n_limit < 0 || totals[1] <= (unsigned long long)n_limit)) *n_written = totals[0];
  if (seg >= n_seg) return;
  const uint32_t c = fb.count[seg];
  if (c == 0) return;
  const uint64_t p = seg / G;
  unsigned long long o = sel_base[p] + seg_sel[seg] - (skip_seg ? sel_base[skip_seg / G] : 0ull);
  unsigned long long cum = reps_base[p] + seg_reps[seg] - (skip_seg ? reps_base[skip_seg / G] : 0ull);
  if (n_limit >= 0 && cum >= (unsigned long long)n_limit) return;
  const long long seg_pos = (long long)(seg * seg_bytes) - (long long)mis;
  for (uint32_t r0 = 0; r0 < c; r0 += 32) {
    const uint32_t r = r0 + lane;
    const uint64_t rr = seg * fb.K + r;
    const uint32_t raw = r < c ? fb.reps[rr] : 0;
    // run records carry 1 + (chosen start - first start) in the chain's output; they count once
    uint32_t reps = ENGINE == FIND_BT_RUN ? (raw != 0 ? 1u : 0u) : raw;
    unsigned long long incl = reps;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += y; }
    const unsigned long long my_cum = cum + incl - reps;
    bool keep = reps != 0;
    if (keep && n_limit >= 0) {
      if (my_cum >= (unsigned long long)n_limit) keep = false;
      else if (my_cum + reps >= (unsigned long long)n_limit) reps = (uint32_t)((unsigned long long)n_limit - my_cum);
    }
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, keep);
    const uint32_t before = __popc(bal & ((1u << lane) - 1u));
    const uint32_t kept = __popc(bal);
    if (keep) {
      const unsigned long long idx = o + before;
      if (n_limit >= 0 && my_cum + reps >= (unsigned long long)n_limit) *n_written = idx + 1;   // the last record kept
      const uint2 k = fb.keys[rr];
      long long* dst = stage + (size_t)before * nc;
      if (ENGINE == FIND_BT_RUN) {
        const long long first = seg_pos + (long long)(int32_t)k.x;
        const long long st = first + (long long)(raw - 1u);       // the start the reference's cursor picked
        dst[0] = st; dst[1] = first + (long long)(k.y >> 12);
        for (int j = 2; j < nc; j += 2) {
          const int32_t a = fb.caps[rr * fb.cw + j - 2], b = fb.caps[rr * fb.cw + j - 1];
          const long long av = ((m.run_start_caps >> j) & 1u) ? st : (a == CAP_ZERO ? 0 : first + a);
          const long long bv = ((m.run_start_caps >> (j + 1)) & 1u) ? st : (b == CAP_ZERO ? 0 : first + b);
          if (av <= bv && bv <= (long long)len) { dst[j] = av; dst[j + 1] = bv; } else { dst[j] = -1; dst[j + 1] = -1; }
        }
      }
      const long long s0 = seg_pos + (long long)k.x, s = s0 + fb.out_base, e = s + (long long)k.y;
      if (ENGINE != FIND_BT_RUN) { dst[0] = s; dst[1] = e; }
      for (int g = 1; ENGINE != FIND_BT_RUN && g < nc / 2; g++) {
        const int32_t a = fb.caps[rr * fb.cw + 2 * g - 2], b = fb.caps[rr * fb.cw + 2 * g - 1];
        if (ENGINE == FIND_TDFA) {
          if (a >= 0) { dst[2 * g] = s + a; dst[2 * g + 1] = s + b; } else { dst[2 * g] = -1; dst[2 * g + 1] = -1; }
        } else {
          const long long av = a == CAP_ZERO ? 0 : s + a, bv = b == CAP_ZERO ? 0 : s + b;
          if (av <= bv && bv <= (long long)len) { dst[2 * g] = av; dst[2 * g + 1] = bv; } else { dst[2 * g] = -1; dst[2 * g + 1] = -1; }
        }
      }
      if (idx < cap_records) out_reps[idx] = reps;
    }
    __syncwarp();
    // contiguous run of kept * nc words starting at record o
    const uint64_t room = o < cap_records ? cap_records - o : 0;
    const uint32_t nrec = (uint32_t)min((uint64_t)kept, room);
    const uint32_t nwords = nrec * (uint32_t)nc;
    long long* gdst = reinterpret_cast<long long*>(out) + o * (uint64_t)nc;
    for (uint32_t w = lane; w < nwords; w += 32) gdst[w] = stage[w];
    __syncwarp();
    o += kept;
    cum += __shfl_sync(0xFFFFFFFFu, incl, 31);
    if (n_limit >= 0 && cum >= (unsigned long long)n_limit) break;
  }
}

}  // namespace rgx
