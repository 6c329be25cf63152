// Cursor replay ("chain") over the record slabs: one lane per part, software-pipelined.
//
// Every lane walks its own part's records in order.  The per-lane stream is contiguous, so a lane
// reads it with 16-byte loads (two {start,len} keys per load) and keeps PF loads in flight ahead of
// the record it is replaying.  A part whose entry cursor is the one it used in the previous pass
// keeps its previous outputs and exits immediately, so after the two warm-up passes only the parts
// that are still being corrected do any work.
#pragma once
#include "kernels_findall2.cuh"

namespace rgx {

constexpr int CHAIN_PF = 4;   // 16-byte loads in flight per lane

// One record of the replay.  `cursor` is relative to the part's first byte (it may be far negative
// when it enters the shard from a predecessor).
constexpr int FIND_BT_RUN = 3;   // backtracking records that stand for a whole run of starts (kernels_btrun.cuh)

template <int ENGINE>
__device__ __forceinline__ uint32_t chain_step(const uint2 k, const long long seg_rel, const long long len_rel, long long& cursor,
                                               uint32_t& nsel, unsigned long long& nreps, int* err) {
  if (k.y == KEY_INVALID) return 0;
  if (ENGINE == FIND_BT_RUN) {
    // key = {first (signed, segment-relative), len << 12 | (last - first)}: every start in [first, last]
    // matches and ends at first + len.  The reference takes the smallest matching start >= cursor and
    // moves the cursor to the match end (find.go:452-457).
    const long long first = seg_rel + (long long)(int32_t)k.x;
    const long long last = first + (long long)(k.y & 0xFFFu);
    if (!(last >= cursor && cursor < len_rel)) return 0;
    const long long s = cursor > first ? cursor : first;
    cursor = first + (long long)(k.y >> 12);
    nsel++;
    nreps++;
    return 1u + (uint32_t)(s - first);
  }
  const long long s = seg_rel + (long long)k.x;
  if (!(s >= cursor && cursor < len_rel)) return 0;
  uint32_t reps;
  if (ENGINE == FIND_TDFA) {
    // offset += len(match) from the SLICE start (compiler.go:630-636): the record at s is returned
    // once per cursor value o, o+L, ... <= s
    const unsigned long long gap = (unsigned long long)(s - cursor);
    const uint32_t L = k.y ? k.y : 1u;
    unsigned long long kk;
    if (gap < (1ull << 22)) {
      // float quotient is within 1 of the true one for gap < 2^22; corrected exactly below
      uint32_t qf = (uint32_t)__fdividef((float)(uint32_t)gap, (float)L);
      uint32_t rem = (uint32_t)gap - qf * L;
      if ((int32_t)rem < 0) { qf--; rem += L; }
      if (rem >= L) { qf++; }
      kk = (unsigned long long)qf + 1ull;
    } else if (gap <= 0xFFFFFFFFull) {
      kk = (unsigned long long)((uint32_t)gap / L) + 1ull;
    } else {
      kk = gap / L + 1ull;
    }
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

struct Chain3Bufs {
  long long* entry_used;          // [n_parts] entry cursor (shard-relative) the part's current outputs were computed from
  unsigned long long* part_sel;   // [n_parts] records kept
  unsigned long long* part_reps;  // [n_parts] matches returned
  uint32_t* seg_sel;              // [n_seg] kept records before this segment, inside its part
  unsigned long long* seg_reps;   // [n_seg] matches returned before this segment, inside its part
  int* changed;                   // [64] changed[pass & 63] set when a part's exit changed in that pass
};

// pass 0: every part replays from the guess "the cursor stands at my first byte" (part 0: the true entry).
// pass j > 0: a part replays from its predecessor's exit unless that is the entry it already used.
// Jacobi iteration: reads the exits of pass j-1 from exit_in, writes exit_out (ping-pong).
template <int ENGINE>
__global__ void __launch_bounds__(64) findall_chain3_kernel(const uint64_t n_seg, const uint32_t seg_bytes, const uint32_t G,
                                                            const uint64_t n_parts, const uint32_t mis, const uint64_t len_in,
                                                            const FindAllBufs fb, const Chain3Bufs cb, const long long* exit_in,
                                                            long long* exit_out, const int pass, int* err) {
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_parts) return;
  const uint64_t seg0 = p * G, seg1 = min(seg0 + G, n_seg);
  const long long part_pos = (long long)(seg0 * seg_bytes) - (long long)mis;   // shard-relative position of the part
  long long entry;
  if (p == 0) entry = fb.entry0;
  else if (pass == 0) entry = part_pos;
  else entry = exit_in[p - 1];
  if (pass > 0 && cb.entry_used[p] == entry) {  // nothing to redo: carry the exit over
    exit_out[p] = exit_in[p];
    return;
  }
  cb.entry_used[p] = entry;
  // `offset < len(input)` refers to the whole logical input
  const long long len_rel = fb.not_last ? (long long)(~0ull >> 2) : (long long)len_in - part_pos;
  long long cursor = entry - part_pos;
  uint32_t nsel = 0;
  unsigned long long nreps = 0;
  for (uint64_t seg = seg0; seg < seg1; seg++) {
    const uint32_t c = fb.count[seg];
    const long long seg_rel = (long long)((seg - seg0) * seg_bytes);
    cb.seg_sel[seg] = nsel;
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
          const uint32_t r0 = chain_step<ENGINE>(make_uint2(cur.x, cur.y), seg_rel, len_rel, cursor, nsel, nreps, err);
          uint32_t r1 = 0;
          if (2 * q + 1 < c) r1 = chain_step<ENGINE>(make_uint2(cur.z, cur.w), seg_rel, len_rel, cursor, nsel, nreps, err);
          rp[q] = make_uint2(r0, r1);
        }
      }
    }
  }
  const long long ex = cursor + part_pos;
  if (pass > 0 && exit_in[p] != ex) cb.changed[pass & 63] = 1;
  exit_out[p] = ex;
  cb.part_sel[p] = nsel;
  cb.part_reps[p] = nreps;
}

// exclusive scan of the per-part counts: one CTA, warp-shuffle scan of 1024-element tiles
__global__ void __launch_bounds__(1024) findall_part_scan2_kernel(const uint64_t n_parts, const unsigned long long* __restrict__ part_sel,
                                                                  const unsigned long long* __restrict__ part_reps,
                                                                  unsigned long long* sel_base, unsigned long long* reps_base,
                                                                  unsigned long long* totals) {
  __shared__ unsigned long long wsum_sel[32], wsum_reps[32];
  __shared__ unsigned long long carry_sel, carry_reps;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { carry_sel = 0; carry_reps = 0; }
  __syncthreads();
  for (uint64_t base = 0; base < n_parts; base += 1024) {
    const uint64_t i = base + threadIdx.x;
    const unsigned long long a = i < n_parts ? part_sel[i] : 0, b = i < n_parts ? part_reps[i] : 0;
    unsigned long long sa = a, sb = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long xa = __shfl_up_sync(0xFFFFFFFFu, sa, o), xb = __shfl_up_sync(0xFFFFFFFFu, sb, o);
      if (lane >= o) { sa += xa; sb += xb; }
    }
    if (lane == 31) { wsum_sel[warp] = sa; wsum_reps[warp] = sb; }
    __syncthreads();
    if (warp == 0) {
      unsigned long long wa = wsum_sel[lane], wb = wsum_reps[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long xa = __shfl_up_sync(0xFFFFFFFFu, wa, o), xb = __shfl_up_sync(0xFFFFFFFFu, wb, o);
        if (lane >= o) { wa += xa; wb += xb; }
      }
      wsum_sel[lane] = wa; wsum_reps[lane] = wb;   // inclusive over warps
    }
    __syncthreads();
    const unsigned long long off_a = carry_sel + (warp ? wsum_sel[warp - 1] : 0), off_b = carry_reps + (warp ? wsum_reps[warp - 1] : 0);
    if (i < n_parts) { sel_base[i] = off_a + sa - a; reps_base[i] = off_b + sb - b; }
    __syncthreads();
    if (threadIdx.x == 0) { carry_sel += wsum_sel[31]; carry_reps += wsum_reps[31]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { totals[0] = carry_sel; totals[1] = carry_reps; }
}

}  // namespace rgx


