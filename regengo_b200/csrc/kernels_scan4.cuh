// findall_scan4_kernel -- the HBM-bound FindAll scan for TDFA patterns with a >= 2-byte literal start
// filter (fourth iteration; same contract as findall_scan_tdfa_kernel in kernels_findall2.cuh).
//
// One warp per 32 KiB segment.
//   FILTER  lanes stream the segment with 16-byte coalesced loads; the loads of the next 2 KiB are in
//           flight while the current 2 KiB are compared (register double buffer).  A two-byte
//           SIMD-in-register compare (3 ALU ops per word and pattern byte, false positives allowed)
//           marks candidate starts; lanes with a hit append its position to the warp's queue at
//           ballot-ranked slots, which keeps the queue sorted without a prefix scan.
//   WALK    32 queued candidates at a time, one TDFA walk per lane, all offsets 32-bit and relative to
//           the segment: a step is byte extraction from a 4-byte word (one aligned load per 4 bytes),
//           one fused shared-memory cell, and three tests.  Tag lists that fire are appended to a
//           per-lane event log; accept lists are logged lazily.
//   REPLAY  with the lanes converged again, tags are rebuilt from the log: apply, in order, every event
//           at or before the last accept, then the accept list in force -- exactly the snapshot
//           tdfa.go:963-975 takes, because a tag value depends only on (list, position).
//   PUBLISH slab entry j of the segment = candidate j ({start_rel, len | KEY_INVALID}, capture offsets).
#pragma once
#include "kernels_findall2.cuh"

namespace rgx {

constexpr uint32_t Q4CAP = 1024;      // candidates per segment before the host falls back to the generic scan
constexpr int LOG4CAP = 12;           // tag events per walk
constexpr int SCAN4_WARPS = 8;

__host__ __device__ inline size_t scan4_extra_words(int ntags) {
  return (size_t)SCAN4_WARPS * ntags * 32 + (size_t)SCAN4_WARPS * LOG4CAP * 32;
}

__global__ void __launch_bounds__(SCAN4_WARPS * 32, 4) findall_scan4_kernel(
    const DevMeta m, const uint32_t* __restrict__ gimg, const uint8_t* __restrict__ buf, const uint64_t len,
    const uint32_t mis, const uint64_t n_seg, const FindAllBufs fb, int* err) {
  extern __shared__ __align__(16) uint32_t smem_all[];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint16_t queue[SCAN4_WARPS][Q4CAP];
  stage_image_tma(smem_all, gimg, m.image_words, &mbar);
  const uint32_t* img = smem_all;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nt = m.t_ntags;
  int32_t* T = reinterpret_cast<int32_t*>(smem_all + m.image_words) + (size_t)warp * nt * 32 + lane;          // tags: T[j*32]
  uint32_t* LG = smem_all + m.image_words + (size_t)SCAN4_WARPS * nt * 32 + (size_t)warp * LOG4CAP * 32 + lane;  // log: LG[e*32]
  uint16_t* q = queue[warp];
  const uint8_t* abuf = buf - mis;                        // 16-byte aligned view of the buffer
  const uint64_t end_a = (uint64_t)mis + fb.cand_len;     // candidate starts are in [mis, end_a)
  const uint64_t load_end = (uint64_t)mis + len;          // bytes exist in [mis, load_end) (shard + halo)
  const uint32_t p0 = (uint32_t)m.prefix_bytes[0] * 0x01010101u;
  const uint32_t p1 = (uint32_t)m.prefix_bytes[1] * 0x01010101u;
  const uint32_t* fast = img + m.off_t_fast;
  const uint32_t* aoff = img + m.off_t_alist_off;
  const uint32_t* alist = img + m.off_t_alist;
  const uint64_t total_warps = (uint64_t)gridDim.x * SCAN4_WARPS;
  const uint32_t lt_mask = (1u << lane) - 1u;
  constexpr uint32_t N_IT = SEG2_BYTES / 512;
  constexpr int U = 4;

  for (uint64_t seg = (uint64_t)blockIdx.x * SCAN4_WARPS + warp; seg < n_seg; seg += total_warps) {
    const uint64_t seg_a = seg * SEG2_BYTES;
    const uint8_t* segp = abuf + seg_a;
    const bool interior = seg_a >= mis && seg_a + SEG2_BYTES <= end_a;
    uint32_t tail = 0;
    bool dense = false;

    // -------- FILTER --------
    auto load16 = [&](uint32_t it) -> uint4 {
      const uint64_t apos = seg_a + (uint64_t)it * 512 + (uint64_t)lane * 16;
      if (interior || (apos >= mis && apos + 16 <= load_end)) return *reinterpret_cast<const uint4*>(abuf + apos);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (apos + 16 > mis && apos < load_end) {
        uint8_t* vb = reinterpret_cast<uint8_t*>(&v);
        for (int j = 0; j < 16; j++) if (apos + j >= mis && apos + j < load_end) vb[j] = abuf[apos + j];
      }
      return v;
    };
    uint4 nxt[U];
#pragma unroll
    for (int u = 0; u < U; u++) nxt[u] = load16(u);
    for (uint32_t it0 = 0; it0 < N_IT; it0 += U) {
      uint4 cur[U];
#pragma unroll
      for (int u = 0; u < U; u++) cur[u] = nxt[u];
      if (it0 + U < N_IT) {
#pragma unroll
        for (int u = 0; u < U; u++) nxt[u] = load16(it0 + U + u);
      }
      uint32_t cc[U][4];
      uint32_t any_all = 0;
#pragma unroll
      for (int u = 0; u < U; u++) {
        const uint32_t t0 = eq_approx(cur[u].x, p1), t1 = eq_approx(cur[u].y, p1), t2 = eq_approx(cur[u].z, p1), t3 = eq_approx(cur[u].w, p1);
        // second-byte flags of the NEXT lane's first word; the last lane cannot see its successor and
        // keeps its 16th byte as a candidate on the first byte alone
        uint32_t t4 = __shfl_down_sync(0xFFFFFFFFu, t0, 1);
        if (lane == 31) t4 = 0x80u;
        cc[u][0] = eq_approx(cur[u].x, p0) & __funnelshift_r(t0, t1, 8);
        cc[u][1] = eq_approx(cur[u].y, p0) & __funnelshift_r(t1, t2, 8);
        cc[u][2] = eq_approx(cur[u].z, p0) & __funnelshift_r(t2, t3, 8);
        cc[u][3] = eq_approx(cur[u].w, p0) & __funnelshift_r(t3, t4, 8);
        any_all |= cc[u][0] | cc[u][1] | cc[u][2] | cc[u][3];
      }
      if (__ballot_sync(0xFFFFFFFFu, any_all != 0)) {
#pragma unroll
        for (int u = 0; u < U; u++) {
          uint32_t mask = 0;
          if (cc[u][0] | cc[u][1] | cc[u][2] | cc[u][3]) {
            mask = gather4(cc[u][0]) | (gather4(cc[u][1]) << 4) | (gather4(cc[u][2]) << 8) | (gather4(cc[u][3]) << 12);
            if (!interior) {
              const uint64_t apos = seg_a + (uint64_t)(it0 + u) * 512 + (uint64_t)lane * 16;
              if (apos >= end_a) mask = 0;
              else {
                if (apos < mis) mask = (mis - apos >= 16) ? 0u : (mask & ~((1u << (uint32_t)(mis - apos)) - 1u));
                if (apos + 16 > end_a) mask &= (1u << (uint32_t)(end_a - apos)) - 1u;
              }
            }
          }
          // ordered append: lane-major order == position order; a lane with several hits takes several rounds
          uint32_t bal = __ballot_sync(0xFFFFFFFFu, mask != 0);
          if (bal) {
            const uint32_t multi = __ballot_sync(0xFFFFFFFFu, (mask & (mask - 1)) != 0);
            if (!multi) {
              const uint32_t slot = tail + __popc(bal & lt_mask);
              if (mask && slot < Q4CAP) q[slot] = (uint16_t)((it0 + u) * 512 + lane * 16 + __ffs(mask) - 1);
              tail += __popc(bal);
            } else {
              // rare: exclusive prefix sum of the per-lane hit counts
              const uint32_t c = __popc(mask);
              uint32_t incl = c;
#pragma unroll
              for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += y; }
              uint32_t w = tail + incl - c;
              while (mask) {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1;
                if (w < Q4CAP) q[w] = (uint16_t)((it0 + u) * 512 + lane * 16 + j);
                w++;
              }
              tail += __shfl_sync(0xFFFFFFFFu, incl, 31);
            }
            if (tail > Q4CAP) dense = true;
          }
        }
      }
    }
    __syncwarp();
    if (dense) { if (lane == 0) atomicOr(err, ERR_DENSE); tail = Q4CAP; }

    // -------- WALK / REPLAY / PUBLISH --------
    // bytes readable from segp, as a 32-bit limit (a walk longer than 4 GiB is out of range)
    const uint64_t avail64 = load_end - seg_a;
    const uint32_t lim = avail64 > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t)avail64;
    const uint32_t lim_eot = avail64 <= 0xFFFFFFF0ull ? lim : 0xFFFFFFFFu;   // ri value that means "the buffer's last byte was just consumed"
    const uint32_t n = tail;
    const uint32_t fast_s = smem_u32(fast);   // shared-window address of the cell table (row = 128 cells = 512 B)
    for (uint32_t base = 0; base < n; base += 32) {
      const uint32_t k = base + lane;
      uint32_t active = k < n ? 1u : 0u;
      const uint32_t srel = active ? q[k] : 0;
      uint32_t ri = srel;                      // next byte to read, relative to segp
      uint32_t row = fast_s + ((uint32_t)m.t_start_any << 9);
      uint32_t acc_ri = 0;                     // ri just after the last accepting step (0: none yet; ri >= 1 there)
      uint32_t pend_al = 0, nlog = 0;
      uint32_t wflags = 0;                     // 1: ran into the end of the buffer, 2: event log overflow
      uint32_t word = 0;
      if (active) {
        if (ri >= lim) { active = 0; wflags = 1; }
        else word = *reinterpret_cast<const uint32_t*>(segp + (ri & ~3u)) >> ((ri & 3u) * 8);
      }
      while (__any_sync(0xFFFFFFFFu, active)) {
        if (active) {
          const uint32_t c = word & 255u;
          uint32_t cell = FAST_NONE;
          if (c < 128) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(cell) : "r"(row + c * 4u));
          if ((cell & 0x3FFu) == FAST_NONE) {
            active = 0;
          } else {
            ri++;
            if (__builtin_expect((cell & 0x000FFC00u) != 0, 0)) {  // a transition tag list fires at position ri - srel
              const uint32_t pos = ri - srel;
              if (pend_al) { if (nlog < LOG4CAP) LG[nlog * 32] = (pend_al << 22) | (acc_ri - srel); nlog++; pend_al = 0; }
              if (nlog < LOG4CAP) LG[nlog * 32] = (((cell >> 10) & 0x3FFu) << 22) | pos;
              nlog++;
              if (pos >= (1u << 22)) wflags |= 2u;
            }
            row = fast_s + ((cell & 0x3FFu) << 9);
            if (cell & 0xC0000000u) {
              if ((cell & 0x40000000u) || ri == lim_eot) {
                const uint32_t aal = (cell >> 20) & 0x3FFu;
                if (__builtin_expect(aal != pend_al, 0)) {
                  if (pend_al) { if (nlog < LOG4CAP) LG[nlog * 32] = (pend_al << 22) | (acc_ri - srel); nlog++; }
                  pend_al = aal;
                }
                acc_ri = ri;
              }
            }
            if (ri >= lim) { active = 0; wflags |= 1u; }
            else { word >>= 8; if ((ri & 3u) == 0) word = *reinterpret_cast<const uint32_t*>(segp + ri); }
          }
        }
      }
      const int32_t match_end = acc_ri ? (int32_t)(acc_ri - srel) : -1;   // relative to the candidate start
      const bool hit_end = (wflags & 1u) != 0, log_ovf = (wflags & 2u) != 0;
      if (hit_end && fb.not_last) atomicOr(err, ERR_HALO);   // ran off the halo: this shard cannot decide the match alone
      if (nlog > LOG4CAP || log_ovf || match_end >= (1 << 22)) atomicOr(err, ERR_DENSE);  // generic scan instead
      // REPLAY
      for (int j = 0; j < nt; j++) T[j * 32] = -1;
      T[0] = 0;
      for (int t = 0; t < m.t_n_init_any; t++) T[img[m.off_t_init + m.t_n_init_begin + t] * 32] = 0;
      const bool matched = k < n && match_end >= 0;
      const uint32_t nmine = matched ? min(nlog, (uint32_t)LOG4CAP) : 0;
      uint32_t nmax = nmine;
#pragma unroll
      for (int o = 16; o; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xFFFFFFFFu, nmax, o));
      for (uint32_t e = 0; e < nmax; e++) {
        if (e < nmine) {
          const uint32_t ev = LG[e * 32];
          const int32_t pos = (int32_t)(ev & 0x3FFFFFu);
          const uint32_t li = ev >> 22;
          if (pos <= match_end)
            for (uint32_t a = aoff[li]; a < aoff[li + 1]; a++) { const uint32_t x = alist[a]; T[(x & 0xFFFFu) * 32] = pos - (int32_t)(x >> 16); }
        }
      }
      if (matched && pend_al)
        for (uint32_t a = aoff[pend_al]; a < aoff[pend_al + 1]; a++) { const uint32_t x = alist[a]; T[(x & 0xFFFFu) * 32] = match_end - (int32_t)(x >> 16); }
      // PUBLISH
      if (k < n) {
        const uint64_t r = seg * fb.K + k;
        if (k < fb.K) {
          if (match_end >= 0) {
            fb.keys[r] = make_uint2(srel, (uint32_t)match_end);
            for (int j = 2; j < nt; j += 2) {
              const int32_t a = T[j * 32];
              int32_t b = T[(j + 1) * 32];
              if (a >= 0 && b < 0) b = match_end;  // unset group end := match end (tdfa.go:1039-1041)
              fb.caps[r * fb.cw + (j - 2)] = a;
              fb.caps[r * fb.cw + (j - 1)] = b;
            }
          } else {
            fb.keys[r] = make_uint2(srel, KEY_INVALID);
          }
        } else {
          atomicOr(err, ERR_SLAB);
        }
      }
      __syncwarp();
    }
    if (lane == 0) fb.count[seg] = min(n, fb.K);
  }
}

}  // namespace rgx
