// Cursor replay ("chain") over the record slabs: one lane per part, software-pipelined.
//
// Every lane walks its own part's records in order.  The per-lane stream is contiguous, so a lane
// reads it with 16-byte loads (two {start,len} keys per load) and keeps PF loads in flight ahead of
// the record it is replaying; a warp-level load instruction therefore touches 32 different sectors,
// but each lane's dependent chain (compare, divide, advance the cursor) never waits on DRAM.
#pragma once
#include "kernels_findall2.cuh"

namespace rgx {

constexpr int CHAIN_PF = 4;   // 16-byte loads in flight per lane

template <int ENGINE>
__device__ __forceinline__ uint32_t chain_step(const uint2 k, const long long seg_pos, const uint64_t len, long long& cursor,
                                               unsigned long long& nsel, unsigned long long& nreps, int* err) {
  if (k.y == KEY_INVALID) return 0;
  const long long s = seg_pos + (long long)k.x;
  if (!(s >= cursor && cursor < (long long)len)) return 0;   // the cursor may sit before the shard (negative)
  uint32_t reps;
  if (ENGINE == FIND_TDFA) {
    // offset += len(match) from the SLICE start (compiler.go:630-636): the record at s is returned
    // once per cursor value o, o+L, ... <= s
    const unsigned long long gap = (unsigned long long)(s - cursor);
    const uint32_t L = k.y ? k.y : 1u;
    unsigned long long kk;
    if (gap <= 0xFFFFFFFFull) kk = (unsigned long long)((uint32_t)gap / L) + 1ull;
    else kk = gap / L + 1ull;
    if (kk > 0xFFFFFFFFull) atomicOr(err, ERR_RANGE);
    reps = (uint32_t)kk;
    cursor += (long long)(kk * L);
  } else {
    // searchStart = captures[1] if it advanced, else searchStart+1 (find.go:452-457)
    reps = 1;
    cursor = k.y ? s + (long long)k.y : s + 1;
  }
  nsel++;
  nreps += reps;
  return reps;
}

template <int ENGINE>
__global__ void __launch_bounds__(64) findall_chain3_kernel(const uint64_t n_seg, const uint32_t seg_bytes, const uint32_t G,
                                                            const uint64_t n_parts, const uint32_t mis, const uint64_t len_in,
                                                            const FindAllBufs fb, const Chain2Bufs cb, const int pass, int* err) {
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_parts) return;
  const uint64_t seg0 = p * G, seg1 = min(seg0 + G, n_seg);
  long long cursor;
  if (p == 0) cursor = fb.entry0;
  else if (pass == 0) cursor = (long long)(seg0 * seg_bytes) - (long long)mis;
  else cursor = cb.exit_prev[p - 1];
  unsigned long long nsel = 0, nreps = 0;
  for (uint64_t seg = seg0; seg < seg1; seg++) {
    const uint32_t c = fb.count[seg];
    const long long seg_pos = (long long)(seg * seg_bytes) - (long long)mis;
    const uint64_t len = fb.not_last ? ~0ull >> 1 : len_in;   // `offset < len(input)` refers to the whole logical input
    cb.seg_sel[seg] = (uint32_t)nsel;
    cb.seg_reps[seg] = nreps;
    const uint4* kp = reinterpret_cast<const uint4*>(fb.keys + seg * fb.K);   // fb.K is even: 16-byte aligned
    uint2* rp = reinterpret_cast<uint2*>(fb.reps + seg * fb.K);
    const uint32_t n_pairs = (c + 1) >> 1;
    uint4 pre[CHAIN_PF];
#pragma unroll
    for (int u = 0; u < CHAIN_PF; u++) pre[u] = (uint32_t)u < n_pairs ? kp[u] : make_uint4(0, KEY_INVALID, 0, KEY_INVALID);
    for (uint32_t q0 = 0; q0 < n_pairs; q0 += CHAIN_PF) {
#pragma unroll
      for (int u = 0; u < CHAIN_PF; u++) {
        const uint32_t q = q0 + u;
        if (q < n_pairs) {
          const uint4 cur = pre[u];
          if (q + CHAIN_PF < n_pairs) pre[u] = kp[q + CHAIN_PF];
          const uint32_t r0 = chain_step<ENGINE>(make_uint2(cur.x, cur.y), seg_pos, len, cursor, nsel, nreps, err);
          uint32_t r1 = 0;
          if (2 * q + 1 < c) r1 = chain_step<ENGINE>(make_uint2(cur.z, cur.w), seg_pos, len, cursor, nsel, nreps, err);
          rp[q] = make_uint2(r0, r1);
        }
      }
    }
  }
  cb.exit_cur[p] = cursor;
  if (pass > 0 && cb.exit_prev[p] != cursor) cb.changed[pass & 63] = 1;
  cb.part_sel[p] = nsel;
  cb.part_reps[p] = nreps;
}

}  // namespace rgx
