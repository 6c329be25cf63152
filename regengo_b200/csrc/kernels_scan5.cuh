// findall_scan5_kernel -- the FindAll scan for TDFA patterns with a >= 2-byte literal start filter
// (one warp per 32 KiB segment; slab entry j of a segment = candidate j: {start_rel, len}, capture offsets).
//
// Its predecessor was issue-bound (0.83 warp instructions per input byte, profiles/r1_b_ncu_scan4_c3.txt); this
// version cuts the instruction count of both halves:
//
//   FILTER  a lane owns 64 contiguous bytes of every 2 KiB block (four 16-byte loads plus the word that
//           follows, issued right after the previous block was compared).  Per word
//           z = (w ^ p0) | (w>>8 ^ p1) | (w>>16 ^ p2) | (w>>24 ^ p3) has a zero byte exactly where the
//           first PLEN <= 4 bytes of the literal prefix start (exact zero-byte test), so every queued
//           candidate is a verified prefix occurrence and its walk starts behind it.  Flags are packed
//           into two words per lane; a warp prefix sum hands out ordered queue slots.
//   WALK    as soon as 32 candidates are queued (the bytes are still in L1/L2), one TDFA walk per lane in
//           two alternating phases:
//             A  skip "boring" bytes -- cells that stay in the same state without tag actions, recognised
//                as cell == selfcell.  Four bytes per iteration from a funnel-shifted register window,
//                four independent shared-memory cell loads, no bookkeeping at all: an accepting state
//                that loops only moves the last-accept position, which phase B recovers from ri.
//             B  one fully general step (end of buffer, byte >= 128, dead cell, tag list, accept list),
//                executed by all lanes together so that the expensive code runs converged.
//   REPLAY / PUBLISH  unchanged from scan4: tags are rebuilt from the per-lane event log, slab entry j of
//           the segment is candidate j.
//
// Exactness of the lazy accept (tdfa.go:937-975): while the state does not change and no transition
// list fires, the reference re-applies the same accept list at every step, so only the last application
// (position ri when the run ends) is observable; phase B sets acc_ri = ri before it looks at the byte
// that ends the run.
#pragma once
#include "kernels_findall2.cuh"

namespace rgx {

constexpr int SCAN5_WARPS = 8;
constexpr uint32_t Q4CAP = 1024;      // candidates per segment before the host falls back to the generic scan
constexpr int LOG4CAP = 12;           // tag events per walk
constexpr uint32_t BLK5 = 2048;          // bytes per filter block (64 per lane)
constexpr uint32_t NOSELF = 0xFFFFFFFFu;
constexpr uint32_t NOCELL = 0xFFFFFFFEu;   // no cell value has transition list 0x3FF and next state 0x3FE

__host__ __device__ inline size_t scan5_extra_words(int ntags) {
  return (size_t)SCAN5_WARPS * ntags * 32 + (size_t)SCAN5_WARPS * (LOG4CAP + 1) * 32;
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

template <int PLEN>
__global__ void __launch_bounds__(SCAN5_WARPS * 32, 4) findall_scan5_kernel(
    const DevMeta m, const uint32_t* __restrict__ gimg, const uint8_t* __restrict__ buf, const uint64_t len,
    const uint32_t mis, const uint64_t n_seg, const FindAllBufs fb, int* err) {
  extern __shared__ __align__(16) uint32_t smem_all[];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint16_t queue[SCAN5_WARPS][Q4CAP];
  stage_image_tma(smem_all, gimg, m.image_words, &mbar);
  const uint32_t* img = smem_all;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nt = m.t_ntags;
  int32_t* T = reinterpret_cast<int32_t*>(smem_all + m.image_words) + (size_t)warp * nt * 32 + lane;          // tags: T[j*32]
  uint32_t* LG = smem_all + m.image_words + (size_t)SCAN5_WARPS * nt * 32 + (size_t)warp * (LOG4CAP + 1) * 32 + lane;  // log: LG[e*32]
  uint16_t* q = queue[warp];
  const uint8_t* abuf = buf - mis;                        // 16-byte aligned view of the buffer
  const uint64_t end_a = (uint64_t)mis + fb.cand_len;     // candidate starts are in [mis, end_a)
  const uint64_t load_end = (uint64_t)mis + len;          // bytes exist in [mis, load_end) (shard + halo)
  const uint32_t p0 = (uint32_t)m.prefix_bytes[0] * 0x01010101u;
  const uint32_t p1 = (uint32_t)m.prefix_bytes[1] * 0x01010101u;
  const uint32_t p2 = (uint32_t)m.prefix_bytes[2] * 0x01010101u;
  const uint32_t p3 = (uint32_t)m.prefix_bytes[3] * 0x01010101u;
  const uint32_t* fast = img + m.off_t_fast;
  const uint32_t* aoff = img + m.off_t_alist_off;
  const uint32_t* alist = img + m.off_t_alist;
  const uint64_t total_warps = (uint64_t)gridDim.x * SCAN5_WARPS;
  constexpr uint32_t N_BLK = SEG2_BYTES / BLK5;
  const uint32_t fast_s = smem_u32(fast);   // shared-window address of the cell table (row = 128 cells = 512 B)
  const uint32_t selftab_s = smem_u32(img + m.off_t_selftab);
  const uint32_t* adesc = img + m.off_t_adesc;
  const uint32_t skip_len = (uint32_t)m.t_skip_len;
  const uint32_t skip_row = fast_s + ((uint32_t)m.t_skip_state << 9);
  const uint32_t skip_self = img[m.off_t_selftab + m.t_skip_state];

  for (uint64_t seg = (uint64_t)blockIdx.x * SCAN5_WARPS + warp; seg < n_seg; seg += total_warps) {
    const uint64_t seg_a = seg * SEG2_BYTES;
    const uint8_t* segp = abuf + seg_a;
    const bool interior = seg_a >= mis && seg_a + SEG2_BYTES <= end_a;
    uint32_t tail = 0, head = 0;
    bool dense = false;
    // bytes readable from segp, as a 32-bit limit (a walk longer than 4 GiB is out of range)
    const uint64_t avail64 = load_end - seg_a;
    const uint32_t lim = avail64 > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t)avail64;
    const uint32_t lim_eot = avail64 <= 0xFFFFFFF0ull ? lim : 0xFFFFFFFFu;   // ri value that means "the buffer's last byte was just consumed"

    // ---------------- WALK / REPLAY / PUBLISH of queue entries [base, min(base + 32, n)) ----------------
    // Queue entry = segment-relative start of a verified prefix occurrence: the walk begins after
    // t_skip_len <= PLEN bytes in t_skip_state with the tag events of those transitions pre-logged.
    auto walk_batch = [&](const uint32_t base, const uint32_t n) {
      const uint32_t k = base + lane;
      uint32_t active = k < n ? 1u : 0u;
      const uint32_t srel = active ? q[k] : 0;
      uint32_t ri = srel + skip_len;           // next byte to read, relative to segp
      uint32_t row = skip_row;
      uint32_t selfcell = skip_self;           // the cell that keeps this state and fires nothing (NOSELF: none)
      uint32_t cur_acc = 0;                    // the current state accepts (not only at end of text)
      uint32_t acc_ri = 0;                     // ri just after the last accepting step (0: none yet; ri >= 1 there)
      uint32_t pend_al = 0;
      uint32_t wflags = 0;                     // 1: ran into the end of the buffer, 2: event log overflow
      for (int e = 0; e < m.t_pre_n; e++) LG[e * 32] = m.t_pre_ev[e];
      uint32_t nlog = (uint32_t)m.t_pre_n;
      // Lanes in a state without a boring cell step (phase B) until every live lane sits in a looping
      // state; then all of them skip their boring runs together (phase A) and take one step out.
      bool exit_step = false;
      uint32_t pre_cell = NOCELL;
      for (;;) {
        const bool do_b = active && (exit_step || selfcell == NOSELF);
        if (!__any_sync(0xFFFFFFFFu, do_b)) {
          if (!__any_sync(0xFFFFFFFFu, active)) break;
          // phase A: boring bytes, eight per iteration (eight independent cell loads in flight); the cell that
          // ends the run is handed to phase B
          if (active) {
            uint32_t a = ri & ~3u;
            if (a + 24 <= lim) {
              const uint32_t sh = (ri & 3u) * 8u;
              uint32_t x0 = *reinterpret_cast<const uint32_t*>(segp + a);
              uint32_t x1 = *reinterpret_cast<const uint32_t*>(segp + a + 4);
              uint32_t x2 = *reinterpret_cast<const uint32_t*>(segp + a + 8);
              uint32_t x3 = *reinterpret_cast<const uint32_t*>(segp + a + 12);
              for (;;) {
                const uint32_t w0 = __funnelshift_r(x0, x1, sh), w1 = __funnelshift_r(x1, x2, sh);
                if ((w0 | w1) & 0x80808080u) break;
                const uint32_t c0 = lds_u32(row + ((w0 << 2) & 0x3FCu));
                const uint32_t c1 = lds_u32(row + ((w0 >> 6) & 0x3FCu));
                const uint32_t c2 = lds_u32(row + ((w0 >> 14) & 0x3FCu));
                const uint32_t c3 = lds_u32(row + ((w0 >> 22) & 0x3FCu));
                const uint32_t c4 = lds_u32(row + ((w1 << 2) & 0x3FCu));
                const uint32_t c5 = lds_u32(row + ((w1 >> 6) & 0x3FCu));
                const uint32_t c6 = lds_u32(row + ((w1 >> 14) & 0x3FCu));
                const uint32_t c7 = lds_u32(row + ((w1 >> 22) & 0x3FCu));
                if (c0 != selfcell) { pre_cell = c0; break; }
                if (c1 != selfcell) { pre_cell = c1; ri += 1; break; }
                if (c2 != selfcell) { pre_cell = c2; ri += 2; break; }
                if (c3 != selfcell) { pre_cell = c3; ri += 3; break; }
                if (c4 != selfcell) { pre_cell = c4; ri += 4; break; }
                if (c5 != selfcell) { pre_cell = c5; ri += 5; break; }
                if (c6 != selfcell) { pre_cell = c6; ri += 6; break; }
                if (c7 != selfcell) { pre_cell = c7; ri += 7; break; }
                ri += 8; a += 8;
                if (a + 24 > lim) break;
                x0 = x2; x1 = x3;
                x2 = *reinterpret_cast<const uint32_t*>(segp + a + 8);
                x3 = *reinterpret_cast<const uint32_t*>(segp + a + 12);
              }
            }
          }
          exit_step = true;
          continue;
        }
        exit_step = false;
        // phase B: one general step
        if (do_b) {
          if (cur_acc) acc_ri = ri;            // the accepting state looped up to here
          if (ri >= lim) { active = 0; wflags |= 1u; }
          else {
            uint32_t cell = pre_cell;            // phase A already looked the byte at ri up (ri < lim there)
            pre_cell = NOCELL;
            if (cell == NOCELL) {
              const uint32_t c = segp[ri];
              cell = FAST_NONE;
              if (c < 128) cell = lds_u32(row + c * 4u);
            }
            if ((cell & 0x3FFu) == FAST_NONE) {
              active = 0;
            } else {
              ri++;
              if ((cell & 0x000FFC00u) != 0) {  // a transition tag list fires at position ri - srel
                const uint32_t pos = ri - srel;
                if (pend_al) { if (nlog < LOG4CAP) LG[nlog * 32] = (pend_al << 22) | (acc_ri - srel); nlog++; pend_al = 0; }
                if (nlog < LOG4CAP) LG[nlog * 32] = (((cell >> 10) & 0x3FFu) << 22) | pos;
                nlog++;
                if (pos >= (1u << 22)) wflags |= 2u;
              }
              row = fast_s + ((cell & 0x3FFu) << 9);
              selfcell = lds_u32(selftab_s + ((cell & 0x3FFu) << 2));
              cur_acc = (cell >> 30) & 1u;
              if (cell & 0xC0000000u) {
                if (cur_acc || ri == lim_eot) {
                  const uint32_t aal = (cell >> 20) & 0x3FFu;
                  if (aal != pend_al) {
                    if (pend_al) { if (nlog < LOG4CAP) LG[nlog * 32] = (pend_al << 22) | (acc_ri - srel); nlog++; }
                    pend_al = aal;
                  }
                  acc_ri = ri;
                }
              }
            }
          }
        }
      }
      const int32_t match_end = acc_ri ? (int32_t)(acc_ri - srel) : -1;   // relative to the candidate start
      const bool hit_end = (wflags & 1u) != 0, log_ovf = (wflags & 2u) != 0;
      if (hit_end && fb.not_last) atomicOr(err, ERR_HALO);   // ran off the halo: this shard cannot decide the match alone
      if (nlog > LOG4CAP || log_ovf || match_end >= (1 << 22)) atomicOr(err, ERR_DENSE);  // generic scan instead
      // REPLAY: apply, in order, every logged list at or before the last accept, then the accept list in force
      for (int j = 0; j < nt; j++) T[j * 32] = -1;
      T[0] = 0;
      for (int t = 0; t < m.t_n_init_any; t++) T[img[m.off_t_init + m.t_n_init_begin + t] * 32] = 0;
      const bool matched = k < n && match_end >= 0;
      uint32_t nmine = matched ? min(nlog, (uint32_t)LOG4CAP) : 0;
      if (matched && pend_al && nmine < LOG4CAP + 1) { LG[nmine * 32] = (pend_al << 22) | (uint32_t)match_end; nmine++; }
      uint32_t nmax = nmine;
#pragma unroll
      for (int o = 16; o; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xFFFFFFFFu, nmax, o));
      for (uint32_t e = 0; e < nmax; e++) {
        if (e < nmine) {
          const uint32_t ev = LG[e * 32];
          const int32_t pos = (int32_t)(ev & 0x3FFFFFu);
          const uint32_t li = ev >> 22;
          if (pos <= match_end) {
            const uint2 ds = *reinterpret_cast<const uint2*>(adesc + 2 * li);
            const uint32_t na = (ds.y >> 16) & 0xFFu;
            if (na > 0) T[(ds.x & 0xFFu) * 32] = pos - (int32_t)((ds.x >> 8) & 0xFFu);
            if (na > 1) T[((ds.x >> 16) & 0xFFu) * 32] = pos - (int32_t)(ds.x >> 24);
            if (na > 2) T[(ds.y & 0xFFu) * 32] = pos - (int32_t)((ds.y >> 8) & 0xFFu);
            if (ds.y >> 24)
              for (uint32_t a = aoff[li]; a < aoff[li + 1]; a++) { const uint32_t x = alist[a]; T[(x & 0xFFFFu) * 32] = pos - (int32_t)(x >> 16); }
          }
        }
      }
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
    };

    // ---------------- FILTER ----------------
    // bytes outside [mis, load_end) read as 0 and never start a candidate: a prefix cut off by the end of
    // the buffer cannot match (every prefix state is non-accepting)
    const bool interior_ld = interior && seg_a + SEG2_BYTES + 4 <= load_end;
    auto load_guarded = [&](const uint64_t apos, const int nb) -> uint4 {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (apos + nb > mis && apos < load_end) {
        uint8_t* vb = reinterpret_cast<uint8_t*>(&v);
        for (int j = 0; j < nb; j++) if (apos + j >= mis && apos + j < load_end) vb[j] = abuf[apos + j];
      }
      return v;
    };
    uint32_t w[17];
    auto load_block = [&](const uint32_t blk) {
      const uint32_t off = blk * BLK5 + lane * 64;
      if (interior_ld) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const uint4 v = *reinterpret_cast<const uint4*>(segp + off + u * 16);
          w[4 * u] = v.x; w[4 * u + 1] = v.y; w[4 * u + 2] = v.z; w[4 * u + 3] = v.w;
        }
        w[16] = *reinterpret_cast<const uint32_t*>(segp + off + 64);
      } else {
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const uint64_t apos = seg_a + off + u * 16;
          const uint4 v = (apos >= mis && apos + 16 <= load_end) ? *reinterpret_cast<const uint4*>(abuf + apos) : load_guarded(apos, 16);
          w[4 * u] = v.x; w[4 * u + 1] = v.y; w[4 * u + 2] = v.z; w[4 * u + 3] = v.w;
        }
        w[16] = load_guarded(seg_a + off + 64, 4).x;
      }
    };
    load_block(0);
    for (uint32_t blk = 0; blk < N_BLK; blk++) {
      // flags -> two words, bit 8*byte + word (word 0..7): 32 bytes each.  The cheap zero-byte test can
      // only err on the byte right above a true zero byte, i.e. when two hits touch; that (practically
      // never) redoes the block with the exact test.
      uint32_t e0 = 0, e1 = 0;
#pragma unroll
      for (int j = 0; j < 16; j++) {
        uint32_t z = (w[j] ^ p0) | (__funnelshift_r(w[j], w[j + 1], 8) ^ p1);
        if (PLEN > 2) z |= __funnelshift_r(w[j], w[j + 1], 16) ^ p2;
        if (PLEN > 3) z |= __funnelshift_r(w[j], w[j + 1], 24) ^ p3;
        const uint32_t d = (z - 0x01010101u) & ~z & 0x80808080u;
        if (j < 8) e0 |= d >> (7 - j); else e1 |= d >> (15 - j);
      }
      if ((e0 & (e0 >> 8)) | (e1 & (e1 >> 8))) {
        e0 = 0; e1 = 0;
#pragma unroll
        for (int j = 0; j < 16; j++) {
          uint32_t z = (w[j] ^ p0) | (__funnelshift_r(w[j], w[j + 1], 8) ^ p1);
          if (PLEN > 2) z |= __funnelshift_r(w[j], w[j + 1], 16) ^ p2;
          if (PLEN > 3) z |= __funnelshift_r(w[j], w[j + 1], 24) ^ p3;
          const uint32_t d = ~(((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z | 0x7F7F7F7Fu);   // 0x80 exactly in the zero bytes
          if (j < 8) e0 |= d >> (7 - j); else e1 |= d >> (15 - j);
        }
      }
      if (blk + 1 < N_BLK) load_block(blk + 1);
      if (!interior && (e0 | e1)) {
        // clip to the candidate range [mis, end_a)
        const uint64_t apos = seg_a + (uint64_t)blk * BLK5 + (uint64_t)lane * 64;
        for (int half = 0; half < 2; half++) {
          uint32_t e = half ? e1 : e0, keep = 0;
          while (e) {
            const uint32_t b = __ffs(e) - 1;
            e &= e - 1;
            const uint64_t ap = apos + half * 32 + 4 * (b & 7u) + (b >> 3);
            if (ap >= mis && ap < end_a) keep |= 1u << b;
          }
          if (half) e1 = keep; else e0 = keep;
        }
      }
      const uint32_t c = __popc(e0) + __popc(e1);
      if (__ballot_sync(0xFFFFFFFFu, c != 0)) {
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += y; }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (c) {
          // this lane's hits, inserted in position order (the packed flags are byte-major)
          const uint32_t wb = tail + incl - c;
          uint32_t pbase = blk * BLK5 + lane * 64;
          uint32_t i = 0, e = e0;
          for (;;) {
            if (!e) { if (!e1) break; e = e1; e1 = 0; pbase += 32; }
            const uint32_t b = __ffs(e) - 1;
            e &= e - 1;
            const uint32_t pos = pbase + 4 * (b & 7u) + (b >> 3);
            uint32_t j = i;
            while (j > 0 && wb + j - 1 < Q4CAP && q[wb + j - 1] > pos) { if (wb + j < Q4CAP) q[wb + j] = q[wb + j - 1]; j--; }
            if (wb + j < Q4CAP) q[wb + j] = (uint16_t)pos;
            i++;
          }
        }
        tail += total;
        if (tail > Q4CAP) { dense = true; tail = Q4CAP; }
        __syncwarp();
      }
      const bool last = blk + 1 == N_BLK;
      while (tail - head >= 32 || (last && head < tail)) { walk_batch(head, min(head + 32, tail)); head = min(head + 32, tail); }
    }
    if (dense && lane == 0) atomicOr(err, ERR_DENSE);
    if (lane == 0) fb.count[seg] = min(tail, fb.K);
  }
}

}  // namespace rgx
