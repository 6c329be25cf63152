// Cursor replay ("chain") over the record slabs.
//
// chain_step<ENGINE> is the reference's cursor rule for one record; chain_replay_part* replay the records of
// one part (G consecutive segments) in order: one lane per part in the throughput passes (16-byte key loads, a
// few in flight), one warp per part in the short worklist passes.  findall_chain4_kernel drives them with
// worklists (see the comment above Chain4Bufs); the part scans turn per-part counts into output positions.
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

// ---------------------------------------------------------------------------------------------------
// chain4: the same replay driven by WORKLISTS, exits updated in place.
//   pass 0       every part replays from the guess "the cursor stands at my first byte" (part 0: true entry)
//   worklist 1   findall_chain4_seed_kernel: the parts whose predecessor's exit differs from that guess
//   pass j >= 1  thread t < |W_j| replays part W_j[t] from its predecessor's current exit; when its own
//                exit moves it appends its successor to W_{j+1}
// Reading a predecessor's exit while that part is being replayed in the same pass is harmless: whichever
// value is seen, a changed exit re-queues the successor, and a part whose entry equals the one it already
// used skips.  When a worklist comes out empty every entry equals its predecessor's exit, which by
// induction from part 0 is the sequential result.  Redo work is compacted: a pass costs what it repairs.
struct Chain4Bufs {
  long long* entry_used;          // [n_parts]
  long long* exitc;               // [n_parts] current exit cursor (shard-relative)
  unsigned long long* part_sel;   // [n_parts] records kept
  unsigned long long* part_reps;  // [n_parts] matches returned
  uint32_t* seg_sel;              // [n_seg]
  unsigned long long* seg_reps;   // [n_seg]
  uint32_t* wl;                   // [2][n_parts] worklists (ping-pong by pass parity)
  uint32_t* wl_count;             // [64] wl_count[j & 63] = |W_j|
};

template <int ENGINE>
__device__ __forceinline__ long long chain_replay_part(const uint64_t p, const long long entry, const uint64_t n_seg,
                                                       const uint32_t seg_bytes, const uint32_t G, const uint32_t mis,
                                                       const uint64_t len_in, const FindAllBufs& fb, const Chain4Bufs& cb, int* err) {
  const uint64_t seg0 = p * G, seg1 = min(seg0 + G, n_seg);
  const long long part_pos = (long long)(seg0 * seg_bytes) - (long long)mis;   // shard-relative position of the part
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
  cb.part_sel[p] = nsel;
  cb.part_reps[p] = nreps;
  return cursor + part_pos;
}

// TDFA cursor rule in 32-bit, branch-free form.  A single lane replays a part, so the replay is one long
// dependency chain and its speed is instructions-per-record: everything that does not depend on the cursor
// (start, length, reciprocal) is computed off the chain, the chain itself is gap -> quotient estimate ->
// exact correction -> new cursor.  Valid while |cursor - part start| < 2^21 on entry and parts of at most 1 MiB (then every gap stays
// below 2^22 and the float quotient is within 1 of the true one); the caller checks that.
__device__ __forceinline__ uint32_t chain_step_tdfa32(const uint32_t kx, const uint32_t ky, const int32_t seg_rel, const int32_t len32,
                                                      int32_t& c, uint32_t& nsel, unsigned long long& nreps) {
  // off the chain: start, length, 2^32 / L as a 32-bit multiplier (approximate reciprocal; L = 1 saturates)
  const int32_t s = seg_rel + (int32_t)kx;
  const uint32_t L = ky ? ky : 1u;
  const uint32_t M = __float2uint_rz(__fdividef(4294967296.0f, (float)L));
  const bool valid = ky != KEY_INVALID;
  // on the chain: gap -> quotient estimate (within 1 for gap < 2^22) -> exact correction -> cursor
  const bool sel = valid && s >= c && c < len32;
  const uint32_t gap = (uint32_t)(s - c);
  uint32_t qf = __umulhi(gap, M);
  const int32_t rem = (int32_t)(gap - qf * L);
  qf += (rem >= (int32_t)L) ? 1u : 0u;
  qf -= (rem < 0) ? 1u : 0u;
  const uint32_t kk = qf + 1u;
  c = sel ? c + (int32_t)(kk * L) : c;
  const uint32_t reps = sel ? kk : 0u;
  nsel += sel ? 1u : 0u;
  nreps += reps;
  return reps;
}

__device__ __forceinline__ long long chain_replay_part_tdfa32(const uint64_t p, const long long entry, const uint64_t n_seg,
                                                              const uint32_t seg_bytes, const uint32_t G, const uint32_t mis,
                                                              const uint64_t len_in, const FindAllBufs& fb, const Chain4Bufs& cb) {
  constexpr int PF = 8;
  const uint64_t seg0 = p * G, seg1 = min(seg0 + G, n_seg);
  const long long part_pos = (long long)(seg0 * seg_bytes) - (long long)mis;
  const long long len_rel = fb.not_last ? (long long)(~0ull >> 2) : (long long)len_in - part_pos;
  const int32_t len32 = len_rel > 0x7FFFFFFFll ? 0x7FFFFFFF : (int32_t)len_rel;
  int32_t c = (int32_t)(entry - part_pos);
  uint32_t nsel = 0;
  unsigned long long nreps = 0;
  const uint4 none = make_uint4(0, KEY_INVALID, 0, KEY_INVALID);
  for (uint64_t seg = seg0; seg < seg1; seg++) {
    const uint32_t cnt = fb.count[seg];
    const int32_t seg_rel = (int32_t)((seg - seg0) * seg_bytes);
    cb.seg_sel[seg] = nsel;
    cb.seg_reps[seg] = nreps;
    const uint4* __restrict__ kp = reinterpret_cast<const uint4*>(fb.keys + seg * fb.K);
    uint2* __restrict__ rp = reinterpret_cast<uint2*>(fb.reps + seg * fb.K);
    const uint32_t n_pairs = (cnt + 1) >> 1;
    uint4 pre[PF];
#pragma unroll
    for (int u = 0; u < PF; u++) pre[u] = (uint32_t)u < n_pairs ? kp[u] : none;
    // the body is branch-free (missing pairs are invalid keys, stores are predicated) so that the compiler
    // can hoist the cursor-independent work of later records above the chain of earlier ones
    for (uint32_t q0 = 0; q0 < n_pairs; q0 += PF) {
#pragma unroll
      for (int u = 0; u < PF; u++) {
        const uint32_t q = q0 + u;
        const uint4 cur = pre[u];
        pre[u] = q + PF < n_pairs ? kp[q + PF] : none;
        const uint32_t r0 = chain_step_tdfa32(cur.x, cur.y, seg_rel, len32, c, nsel, nreps);
        const uint32_t r1 = chain_step_tdfa32(cur.z, 2 * q + 1 < cnt ? cur.w : KEY_INVALID, seg_rel, len32, c, nsel, nreps);
        if (q < n_pairs) rp[q] = make_uint2(r0, r1);
      }
    }
  }
  cb.part_sel[p] = nsel;
  cb.part_reps[p] = nreps;
  return (long long)c + part_pos;
}

// The same replay by a whole WARP, for the short worklist passes whose cost is one part's latency: lanes load
// 32 keys at once (coalesced) and prepare start / length / reciprocal in parallel; the cursor chain itself runs
// redundantly in every lane over shuffled operands (about 12 issue slots per record instead of ~35 dependent
// ones), and lane j keeps record j's result so that the repeat counts are stored coalesced.
__device__ __forceinline__ long long chain_replay_part_tdfa32_warp(const uint64_t p, const long long entry, const uint64_t n_seg,
                                                                   const uint32_t seg_bytes, const uint32_t G, const uint32_t mis,
                                                                   const uint64_t len_in, const FindAllBufs& fb, const Chain4Bufs& cb) {
  const int lane = threadIdx.x & 31;
  const uint64_t seg0 = p * G, seg1 = min(seg0 + G, n_seg);
  const long long part_pos = (long long)(seg0 * seg_bytes) - (long long)mis;
  const long long len_rel = fb.not_last ? (long long)(~0ull >> 2) : (long long)len_in - part_pos;
  const int32_t len32 = len_rel > 0x7FFFFFFFll ? 0x7FFFFFFF : (int32_t)len_rel;
  int32_t c = (int32_t)(entry - part_pos);
  uint32_t nsel = 0;
  unsigned long long nreps = 0;
  for (uint64_t seg = seg0; seg < seg1; seg++) {
    const uint32_t cnt = fb.count[seg];
    const int32_t seg_rel = (int32_t)((seg - seg0) * seg_bytes);
    if (lane == 0) { cb.seg_sel[seg] = nsel; cb.seg_reps[seg] = nreps; }
    const uint2* kp = fb.keys + seg * fb.K;
    uint32_t* rp = fb.reps + seg * fb.K;
    uint2 nxt = (uint32_t)lane < cnt ? kp[lane] : make_uint2(0, KEY_INVALID);
    for (uint32_t base = 0; base < cnt; base += 32) {
      const uint2 k = nxt;
      if (base + 32 < cnt) nxt = base + 32 + lane < cnt ? kp[base + 32 + lane] : make_uint2(0, KEY_INVALID);
      const int32_t my_s = seg_rel + (int32_t)k.x;
      const uint32_t my_valid = k.y != KEY_INVALID ? 1u : 0u;
      const uint32_t my_L = (k.y && my_valid) ? k.y : 1u;
      const uint32_t my_M = __float2uint_rz(__fdividef(4294967296.0f, (float)my_L));
      uint32_t my_reps = 0;
#pragma unroll
      for (int j = 0; j < 32; j++) {
        const int32_t s = __shfl_sync(0xFFFFFFFFu, my_s, j);
        const uint32_t L = __shfl_sync(0xFFFFFFFFu, my_L, j);
        const uint32_t M = __shfl_sync(0xFFFFFFFFu, my_M, j);
        const uint32_t valid = __shfl_sync(0xFFFFFFFFu, my_valid, j);
        const bool sel = valid && s >= c && c < len32;
        const uint32_t gap = (uint32_t)(s - c);
        uint32_t qf = __umulhi(gap, M);
        const int32_t rem = (int32_t)(gap - qf * L);
        qf += (rem >= (int32_t)L) ? 1u : 0u;
        qf -= (rem < 0) ? 1u : 0u;
        const uint32_t kk = qf + 1u;
        c = sel ? c + (int32_t)(kk * L) : c;
        const uint32_t reps = sel ? kk : 0u;
        nsel += sel ? 1u : 0u;
        nreps += reps;
        if (lane == j) my_reps = reps;
      }
      if (base + lane < cnt) rp[base + lane] = my_reps;
    }
  }
  if (lane == 0) { cb.part_sel[p] = nsel; cb.part_reps[p] = nreps; }
  return (long long)c + part_pos;
}

template <int ENGINE>
__global__ void __launch_bounds__(64) findall_chain4_kernel(const uint64_t n_seg, const uint32_t seg_bytes, const uint32_t G,
                                                            const uint64_t n_parts, const uint32_t mis, const uint64_t len_in,
                                                            const FindAllBufs fb, const Chain4Bufs cb, const int pass, int* err) {
  // a failed scan (dense input, slab overflow, ...) left incomplete slabs behind: the host discards this attempt, and
  // nothing downstream may follow positions read from entries that were never written
  if (*(volatile int*)err) return;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t p;
  long long entry;
  if (ENGINE == FIND_TDFA && pass > 0) {
    // short worklists: one warp per part (latency mode)
    const uint32_t cnt = cb.wl_count[pass & 63];
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    if (cnt <= n_warps && (uint64_t)G * seg_bytes <= (1u << 20)) {
      const uint64_t wid = t >> 5;
      if (wid >= cnt) return;
      p = cb.wl[(size_t)(pass & 1) * n_parts + wid];
      entry = *reinterpret_cast<volatile long long*>(cb.exitc + (p - 1));
      const long long used = cb.entry_used[p];
      __syncwarp();
      if (entry == used) return;
      const long long rel = entry - ((long long)(p * G * seg_bytes) - (long long)mis);
      long long ex;
      if (rel > -(1ll << 21) && rel < (1ll << 21)) {
        ex = chain_replay_part_tdfa32_warp(p, entry, n_seg, seg_bytes, G, mis, len_in, fb, cb);
      } else {
        ex = 0;
        if ((threadIdx.x & 31) == 0) ex = chain_replay_part<ENGINE>(p, entry, n_seg, seg_bytes, G, mis, len_in, fb, cb, err);
        ex = __shfl_sync(0xFFFFFFFFu, ex, 0);
      }
      if ((threadIdx.x & 31) == 0) {
        cb.entry_used[p] = entry;
        if (cb.exitc[p] != ex) {
          *reinterpret_cast<volatile long long*>(cb.exitc + p) = ex;
          if (p + 1 < n_parts) {
            const uint32_t slot = atomicAdd(&cb.wl_count[(pass + 1) & 63], 1u);
            cb.wl[(size_t)((pass + 1) & 1) * n_parts + slot] = (uint32_t)(p + 1);
          }
        }
      }
      return;
    }
  }
  if (pass == 0) {
    if (t >= n_parts) return;
    p = t;
    entry = p == 0 ? fb.entry0 : (long long)(p * G * seg_bytes) - (long long)mis;
  } else {
    if (t >= cb.wl_count[pass & 63]) return;
    p = cb.wl[(size_t)(pass & 1) * n_parts + t];
    entry = *reinterpret_cast<volatile long long*>(cb.exitc + (p - 1));
    if (entry == cb.entry_used[p]) return;
  }
  cb.entry_used[p] = entry;
  long long ex;
  const long long rel = entry - ((long long)(p * G * seg_bytes) - (long long)mis);
  if (ENGINE == FIND_TDFA && rel > -(1ll << 21) && rel < (1ll << 21) && (uint64_t)G * seg_bytes <= (1u << 20))
    ex = chain_replay_part_tdfa32(p, entry, n_seg, seg_bytes, G, mis, len_in, fb, cb);
  else
    ex = chain_replay_part<ENGINE>(p, entry, n_seg, seg_bytes, G, mis, len_in, fb, cb, err);
  if (pass == 0) { cb.exitc[p] = ex; return; }
  if (cb.exitc[p] != ex) {
    *reinterpret_cast<volatile long long*>(cb.exitc + p) = ex;
    if (p + 1 < n_parts) {
      const uint32_t slot = atomicAdd(&cb.wl_count[(pass + 1) & 63], 1u);
      cb.wl[(size_t)((pass + 1) & 1) * n_parts + slot] = (uint32_t)(p + 1);
    }
  }
}

// W_1 = parts whose pass-0 guess was wrong
__global__ void __launch_bounds__(256) findall_chain4_seed_kernel(const uint64_t n_parts, const Chain4Bufs cb) {
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool need = p >= 1 && p < n_parts && cb.exitc[p - 1] != cb.entry_used[p];
  const uint32_t bal = __ballot_sync(0xFFFFFFFFu, need);
  if (!bal) return;
  const int lane = threadIdx.x & 31;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(&cb.wl_count[1], (uint32_t)__popc(bal));
  base = __shfl_sync(0xFFFFFFFFu, base, 0);
  if (need) cb.wl[n_parts + base + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)p;
}

// Exclusive scan of the per-part counts in two launches: tile sums (1024 parts per CTA), then every CTA
// adds the sums of the tiles before it and scans its own tile.
__global__ void __launch_bounds__(1024) findall_part_tilesum_kernel(const uint64_t n_parts, const unsigned long long* __restrict__ part_sel,
                                                                    const unsigned long long* __restrict__ part_reps,
                                                                    unsigned long long* tile_sel, unsigned long long* tile_reps) {
  __shared__ unsigned long long ws[32], wr[32];
  const uint64_t i = (uint64_t)blockIdx.x * 1024 + threadIdx.x;
  unsigned long long a = i < n_parts ? part_sel[i] : 0, b = i < n_parts ? part_reps[i] : 0;
#pragma unroll
  for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(0xFFFFFFFFu, a, o); b += __shfl_xor_sync(0xFFFFFFFFu, b, o); }
  if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5] = a; wr[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x < 32) {
    a = ws[threadIdx.x]; b = wr[threadIdx.x];
#pragma unroll
    for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(0xFFFFFFFFu, a, o); b += __shfl_xor_sync(0xFFFFFFFFu, b, o); }
    if (threadIdx.x == 0) { tile_sel[blockIdx.x] = a; tile_reps[blockIdx.x] = b; }
  }
}

__global__ void __launch_bounds__(1024) findall_part_scan3_kernel(const uint64_t n_parts, const unsigned long long* __restrict__ part_sel,
                                                                  const unsigned long long* __restrict__ part_reps,
                                                                  const unsigned long long* __restrict__ tile_sel,
                                                                  const unsigned long long* __restrict__ tile_reps,
                                                                  unsigned long long* sel_base, unsigned long long* reps_base,
                                                                  unsigned long long* totals) {
  __shared__ unsigned long long wsum_sel[32], wsum_reps[32];
  __shared__ unsigned long long carry_sel, carry_reps;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // sums of the tiles before this one (and, in the last CTA, the grand totals)
  {
    unsigned long long a = 0, b = 0;
    for (uint32_t j = threadIdx.x; j < blockIdx.x; j += 1024) { a += tile_sel[j]; b += tile_reps[j]; }
#pragma unroll
    for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(0xFFFFFFFFu, a, o); b += __shfl_xor_sync(0xFFFFFFFFu, b, o); }
    if (lane == 0) { wsum_sel[warp] = a; wsum_reps[warp] = b; }
    __syncthreads();
    if (warp == 0) {
      a = wsum_sel[lane]; b = wsum_reps[lane];
#pragma unroll
      for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(0xFFFFFFFFu, a, o); b += __shfl_xor_sync(0xFFFFFFFFu, b, o); }
      if (lane == 0) { carry_sel = a; carry_reps = b; }
    }
    __syncthreads();
  }
  const uint64_t i = (uint64_t)blockIdx.x * 1024 + threadIdx.x;
  const unsigned long long a = i < n_parts ? part_sel[i] : 0, b = i < n_parts ? part_reps[i] : 0;
  unsigned long long sa = a, sb = b;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long xa = __shfl_up_sync(0xFFFFFFFFu, sa, o), xb = __shfl_up_sync(0xFFFFFFFFu, sb, o);
    if (lane >= o) { sa += xa; sb += xb; }
  }
  __syncthreads();
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
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 1023) { totals[0] = carry_sel + wsum_sel[31]; totals[1] = carry_reps + wsum_reps[31]; }
}

}  // namespace rgx


