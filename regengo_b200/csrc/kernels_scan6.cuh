// findall_scan6_kernel -- the FindAll scan for TDFA patterns with a literal start (one warp per SPAN of U
// consecutive 32 KiB segments; slab entry j of a segment = candidate j: {start_rel, len}, capture offsets).
//
// scan5 (round 1) walked candidates in batches of 32 and was issue-bound with a third of the lanes busy:
// lanes of a batch finish at different times, and the general step carried the whole accept/tag bookkeeping.
// This kernel keeps the lanes busy and makes the steps cheap:
//
//   FILTER    a lane owns 64 contiguous bytes of every 2 KiB block (four 16-byte loads, the next block's issued
//             before this one is compared).  Per word, z = (w ^ p) | (w >> 8d ^ q) has a zero byte exactly where
//             byte p is followed, d bytes later, by byte q -- the first and the last of the first four bytes of the
//             literal prefix.  Both are NECESSARY for a match, so no start is lost; the walk checks the rest.
//             Hits are queued in position order (ballot rank; the general ordered insert only when a lane has two).
//   WALK      a lane is a walker: it takes the next queued candidate the moment its previous walk ends (refill),
//             so walks of different lengths overlap instead of waiting for each other.  One iteration =
//               burst   up to 4 CHEAP steps per lane from a register window: cell = row[byte]; a cheap cell IS the
//                       shared-memory address of the next row (relocated once per CTA), so a step is byte
//                       extract, address, LDS, sign test.
//               event   one step for the lanes that met a non-cheap cell: append {descriptor, position} to the
//                       candidate's log, follow the descriptor's row.  A dead cell (or the end of the buffer)
//                       ends the walk: the lane writes a one-word header and goes idle.
//             Nothing else happens during a walk: which events count, where the match ends and what the tags are is
//             worked out afterwards from the log.
//   FINALIZE  logs live in a ring of 64 slots indexed by candidate number; as soon as 32 CONSECUTIVE candidates
//             are finished one pass replays their logs -- one candidate per lane, all lanes busy -- and
//             publishes the records.
//
// Exactness of reading the match off the log (tdfa.go:929-983).  Between two logged events the walk takes only
// cheap steps; by construction (device_program.cu) those fire no transition list and either stay in the state
// or move between states that never accept.  So after event e the automaton stays in the event's next state S_e
// for as long as it accepts, and the reference re-applies S_e's accept list at every step of that run: only the
// last application, at the position where the next event (or the end of the walk) begins, is observable.  The
// last accepting step of the walk is therefore the end of the run of the LAST event whose next state accepts
// (`lastacc`), and the tags at that moment are: for each event up to it, its transition list at its own position,
// then (if S_e accepts) S_e's accept list at the end of its run.
#pragma once
#include "kernels_findall2.cuh"

namespace rgx {

constexpr int S6_SLOTS = 64;             // candidates in flight per warp (power of two)
constexpr int S6_LOGCAP = 11;            // events per walk
constexpr int S6_SLOTW = 2 + S6_LOGCAP;  // header, start, events (odd: consecutive slots fall into different banks)
constexpr uint32_t S6_QCAP = 128;        // queued candidate starts per warp (power of two)
constexpr uint32_t S6_BLK = 2048;        // bytes per filter block (64 per lane)
constexpr uint32_t S6_NOEV = 0;          // no pending cell (a cheap cell is a row address, never 0)
constexpr uint32_t S6_SLOW = 1;          // pending: too close to the end of the buffer for the register window
constexpr int S6_MAXU = 4;               // segments per span, at most

__host__ __device__ inline size_t scan6_warp_words(int ntags) {
  return (size_t)S6_SLOTS * S6_SLOTW + S6_QCAP + (size_t)ntags * 32 + 8;
}

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// D: distance of the second filter byte (0: single-byte filter).  WARPS spans per CTA, one per warp.
template <int D, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) findall_scan6_kernel(
    const DevMeta m, const uint32_t* __restrict__ gimg, const uint8_t* __restrict__ buf, const uint64_t len,
    const uint32_t mis, const uint64_t n_seg, const uint32_t U, const FindAllBufs fb, int* err) {
  extern __shared__ __align__(16) uint32_t smem_all[];
  __shared__ __align__(8) unsigned long long mbar;
  stage_image_tma(smem_all, gimg + m.w6_off, m.w6_words, &mbar);
  const uint32_t rows_s = smem_u32(smem_all);
  // relocation: cheap cells and descriptor rows become absolute shared-memory addresses
  {
    const uint32_t n_cells = (uint32_t)m.t_ns * 256u;
    for (uint32_t i = threadIdx.x; i < n_cells; i += WARPS * 32) {
      const uint32_t c = smem_all[i];
      if (!(c & S6_EVBIT)) smem_all[i] = c + rows_s;
    }
    for (uint32_t i = 1 + threadIdx.x; i < (uint32_t)m.w6_ndesc; i += WARPS * 32) smem_all[m.w6_desc + 2 * i] += rows_s;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int nt = m.t_ntags;
  uint32_t* wbase = smem_all + m.w6_words + (size_t)warp * scan6_warp_words(nt);
  const uint32_t slots_s = smem_u32(wbase);                               // S6_SLOTS x S6_SLOTW words
  uint32_t* Q = wbase + S6_SLOTS * S6_SLOTW;                               // S6_QCAP starts (span-relative)
  int32_t* T = reinterpret_cast<int32_t*>(Q + S6_QCAP) + lane;             // tags of the replay: T[j * 32]
  uint32_t* kb = Q + S6_QCAP + (size_t)nt * 32;                            // kb[s] = candidates of the span before its segment s
  const uint32_t desc_s = rows_s + m.w6_desc * 4u;
  const uint32_t* adesc = smem_all + m.w6_adesc;
  const uint32_t* aoff = smem_all + m.w6_aoff;
  const uint32_t* alist = smem_all + m.w6_alist;
  const uint32_t* initt = smem_all + m.w6_init;
  const uint32_t start_row = rows_s + (uint32_t)m.t_start_any * 1024u;

  const uint64_t span = (uint64_t)blockIdx.x * WARPS + warp;
  const uint64_t seg0 = span * U;
  if (seg0 >= n_seg) return;
  const uint32_t n_sseg = (uint32_t)min((uint64_t)U, n_seg - seg0);        // segments of this span
  const uint8_t* abuf = buf - mis;                        // 16-byte aligned view of the buffer
  const uint64_t end_a = (uint64_t)mis + fb.cand_len;     // candidate starts are in [mis, end_a)
  const uint64_t load_end = (uint64_t)mis + len;          // bytes exist in [mis, load_end) (shard + halo)
  const uint64_t span_a = seg0 * SEG2_BYTES;
  const uint8_t* sp = abuf + span_a;
  const uint32_t span_bytes = n_sseg * SEG2_BYTES;
  const bool interior = span_a >= mis && span_a + span_bytes <= end_a;
  const bool interior_ld = interior && span_a + span_bytes + 4 <= load_end;
  // bytes readable from sp, as a 32-bit limit (a walk longer than 4 GiB is out of range)
  const uint64_t avail64 = load_end - span_a;
  const uint32_t lim = avail64 > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t)avail64;
  const uint32_t lim_eot = avail64 <= 0xFFFFFFF0ull ? lim : 0xFFFFFFFFu;   // ri value that means "the buffer's last byte was just consumed"
  const uint32_t pv = (uint32_t)m.w6_p * 0x01010101u, qv = (uint32_t)m.w6_q * 0x01010101u;

  // ---- walker state (one per lane) ----
  uint32_t active = 0, k = 0, srel = 0, ri = 0, row = 0, nlog = 0, lastacc = 0, pend = S6_NOEV, x0 = 0, x1 = 0, slot_s = 0;
  // ---- warp-uniform bookkeeping ----
  uint32_t q_head = 0, q_tail = 0;      // queue ring counters
  uint32_t next_k = 0, fin_base = 0;    // next candidate number to hand out; first candidate not yet finalized
  uint32_t done_lo = 0, done_hi = 0;    // finished flags of candidates fin_base + 0..31 / + 32..63
  bool dense = false;

  // FINALIZE candidates [base, base + cnt): replay the logs, publish the records
  auto finalize = [&](const uint32_t base, const uint32_t cnt) {
    const uint32_t kk = base + lane;
    const bool valid = (uint32_t)lane < cnt;
    const uint32_t ss = slots_s + (kk & (S6_SLOTS - 1)) * (S6_SLOTW * 4u);
    const uint32_t hdr = valid ? lds32(ss) : 0u;
    const uint32_t st = valid ? lds32(ss + 4) : 0u;
    const uint32_t nl = (hdr >> 22) & 15u, end_rel = hdr & 0x3FFFFFu;
    uint32_t la = (hdr >> 26) & 15u;
    if (hdr >> 31) atomicOr(err, ERR_DENSE);       // log overflow or a walk of 4 MiB: the generic scan decides
    if (((hdr >> 30) & 1u) && nl) {                // the walk consumed the last byte of the input: acceptStatesEOT counts
      const uint32_t d = lds32(desc_s + (lds32(ss + 8 + (nl - 1) * 4) >> 22) * 8u + 4u);
      if (d & S6_ACC_EOT) la = nl;
    }
    const bool matched = valid && la > 0;
    const int32_t match_end = !matched ? -1 : (la < nl ? (int32_t)(lds32(ss + 8 + la * 4) & 0x3FFFFFu) : (int32_t)end_rel);
    for (int j = 0; j < nt; j++) T[j * 32] = -1;
    T[0] = 0;
    for (int t = 0; t < m.t_n_init_any; t++) T[initt[t] * 32] = 0;
    auto apply = [&](const uint32_t li, const int32_t pos) {
      const uint2 ds = *reinterpret_cast<const uint2*>(adesc + 2 * li);
      const uint32_t na = (ds.y >> 16) & 0xFFu;
      if (na > 0) T[(ds.x & 0xFFu) * 32] = pos - (int32_t)((ds.x >> 8) & 0xFFu);
      if (na > 1) T[((ds.x >> 16) & 0xFFu) * 32] = pos - (int32_t)(ds.x >> 24);
      if (na > 2) T[(ds.y & 0xFFu) * 32] = pos - (int32_t)((ds.y >> 8) & 0xFFu);
      if (ds.y >> 24)
        for (uint32_t a = aoff[li]; a < aoff[li + 1]; a++) { const uint32_t x = alist[a]; T[(x & 0xFFFFu) * 32] = pos - (int32_t)(x >> 16); }
    };
    uint32_t nmax = matched ? la : 0u;
    nmax = __reduce_max_sync(0xFFFFFFFFu, nmax);
    for (uint32_t e = 0; e < nmax; e++) {
      if (matched && e < la) {
        const uint32_t ev = lds32(ss + 8 + e * 4);
        const uint32_t d = lds32(desc_s + (ev >> 22) * 8u + 4u);
        const uint32_t tl = d & 0x3FFu, al = (d >> 10) & 0x3FFu;
        if (tl) apply(tl, (int32_t)(ev & 0x3FFFFFu) + 1);
        if (al && ((d & S6_ACC) || e + 1 == la)) {
          const int32_t run_end = e + 1 < nl ? (int32_t)(lds32(ss + 8 + (e + 1) * 4) & 0x3FFFFFu) : (int32_t)end_rel;
          apply(al, run_end);
        }
      }
    }
    if (valid) {
      const uint32_t sseg = st >> 15;                       // segment of the span the candidate starts in
      const uint32_t j = kk - kb[sseg];
      const uint64_t r = (seg0 + sseg) * fb.K + j;
      if (j < fb.K) {
        if (matched) {
          fb.keys[r] = make_uint2(st & (SEG2_BYTES - 1), (uint32_t)match_end);
          for (int q = 2; q < nt; q += 2) {
            const int32_t a = T[q * 32];
            int32_t b = T[(q + 1) * 32];
            if (a >= 0 && b < 0) b = match_end;  // unset group end := match end (tdfa.go:1039-1041)
            fb.caps[r * fb.cw + (q - 2)] = a;
            fb.caps[r * fb.cw + (q - 1)] = b;
          }
        } else {
          fb.keys[r] = make_uint2(st & (SEG2_BYTES - 1), KEY_INVALID);
        }
      } else {
        atomicOr(err, ERR_SLAB);
      }
    }
    __syncwarp();
  };

  // WALK: iterate while candidates are queued (drain: until every walk has ended and every record is out)
  auto walk_run = [&](const bool drain) {
    for (;;) {
      const uint32_t idle_m = __ballot_sync(0xFFFFFFFFu, !active);
      const uint32_t qn = q_tail - q_head;
      if (qn == 0) {
        if (!drain || idle_m == 0xFFFFFFFFu) break;
      } else if (idle_m) {
        // refill: the idle lanes take the next candidates, in lane order (a slot must be free: fin_base + 64 > k)
        const uint32_t navail = min(qn, fin_base + S6_SLOTS - next_k);
        const uint32_t rank = __popc(idle_m & lt_mask);
        if (!active && rank < navail) {
          k = next_k + rank;
          srel = Q[(q_head + rank) & (S6_QCAP - 1)];
          slot_s = slots_s + (k & (S6_SLOTS - 1)) * (S6_SLOTW * 4u);
          sts32(slot_s + 4, srel);
          ri = srel; row = start_row; nlog = 0; lastacc = 0; pend = S6_NOEV; active = 1;
          const uint32_t a = ri & ~3u;
          if (a + 8 <= lim) { x0 = *reinterpret_cast<const uint32_t*>(sp + a); x1 = *reinterpret_cast<const uint32_t*>(sp + a + 4); }
        }
        const uint32_t taken = min((uint32_t)__popc(idle_m), navail);
        q_head += taken; next_k += taken;
      }
      // burst: up to four cheap steps
      if (active && pend == S6_NOEV) {
        uint32_t a = ri & ~3u;
        if (a + 8 > lim) pend = S6_SLOW;
        else {
          const uint32_t win = __funnelshift_r(x0, x1, (ri & 3u) * 8u);
          uint32_t c = lds32(row + ((win << 2) & 0x3FCu));
          if (c & S6_EVBIT) pend = c;
          else {
            row = c; c = lds32(row + ((win >> 6) & 0x3FCu));
            if (c & S6_EVBIT) { pend = c; ri += 1; }
            else {
              row = c; c = lds32(row + ((win >> 14) & 0x3FCu));
              if (c & S6_EVBIT) { pend = c; ri += 2; }
              else {
                row = c; c = lds32(row + ((win >> 22) & 0x3FCu));
                if (c & S6_EVBIT) { pend = c; ri += 3; }
                else { row = c; ri += 4; }
              }
            }
          }
          // keep the window on ri: x0 = word at ri & ~3, x1 = the next one
          if ((ri & ~3u) != a) { a += 4; x0 = x1; if (a + 8 <= lim) x1 = *reinterpret_cast<const uint32_t*>(sp + a + 4); }
        }
      }
      // event: one step for the lanes that stopped at a non-cheap cell
      bool fin_now = false;
      if (active && pend != S6_NOEV) {
        uint32_t cell = pend;
        pend = S6_NOEV;
        bool eob = false;
        if (cell == S6_SLOW) {
          if (ri >= lim) { eob = true; cell = S6_EVBIT; }
          else {
            cell = lds32(row + (uint32_t)sp[ri] * 4u);
            if (!(cell & S6_EVBIT)) { row = cell; ri++; cell = S6_NOEV; }
          }
        }
        if (cell != S6_NOEV) {
          const uint32_t idx = cell & 0xFFFFu;
          if (idx == 0) {
            // dead cell or end of the buffer: the walk is over
            const uint32_t end_rel = ri - srel;
            uint32_t hdr = (end_rel & 0x3FFFFFu) | (min(nlog, 15u) << 22) | (lastacc << 26);
            if (eob && ri == lim_eot) hdr |= 1u << 30;
            if (end_rel >= (1u << 22) || nlog > (uint32_t)S6_LOGCAP) hdr |= 1u << 31;
            if (eob && fb.not_last) atomicOr(err, ERR_HALO);   // ran off the halo: this shard cannot decide the match alone
            sts32(slot_s, hdr);
            active = 0; fin_now = true;
          } else {
            const uint2 d = lds64(desc_s + idx * 8u);
            if (nlog < (uint32_t)S6_LOGCAP) sts32(slot_s + 8 + nlog * 4, (idx << 22) | ((ri - srel) & 0x3FFFFFu));
            nlog++;
            if (d.y & S6_ACC) lastacc = min(nlog, 15u);
            row = d.x;
            ri++;
            if ((ri & 3u) == 0) { x0 = x1; if (ri + 8 <= lim) x1 = *reinterpret_cast<const uint32_t*>(sp + ri + 4); }
          }
        }
      }
      const uint32_t fm = __ballot_sync(0xFFFFFFFFu, fin_now);
      if (fm) {
        const uint32_t b = k - fin_base;
        done_lo |= __reduce_or_sync(0xFFFFFFFFu, (fin_now && b < 32) ? (1u << b) : 0u);
        done_hi |= __reduce_or_sync(0xFFFFFFFFu, (fin_now && b >= 32) ? (1u << (b - 32)) : 0u);
        while (done_lo == 0xFFFFFFFFu) {
          finalize(fin_base, 32);
          fin_base += 32; done_lo = done_hi; done_hi = 0;
        }
      }
    }
    if (drain) {
      while (fin_base < next_k) {
        finalize(fin_base, min(32u, next_k - fin_base));
        fin_base += 32; done_lo = done_hi; done_hi = 0;
      }
    }
  };

  // ---------------- FILTER ----------------
  // bytes outside [mis, load_end) read as 0 and never start a candidate: a prefix cut off by the end of
  // the buffer cannot match (every prefix state is non-accepting)
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
    const uint32_t off = blk * S6_BLK + lane * 64;
    if (interior_ld) {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint4 v = *reinterpret_cast<const uint4*>(sp + off + u * 16);
        w[4 * u] = v.x; w[4 * u + 1] = v.y; w[4 * u + 2] = v.z; w[4 * u + 3] = v.w;
      }
      if (D > 0) w[16] = *reinterpret_cast<const uint32_t*>(sp + off + 64);
    } else {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint64_t apos = span_a + off + u * 16;
        const uint4 v = (apos >= mis && apos + 16 <= load_end) ? *reinterpret_cast<const uint4*>(abuf + apos) : load_guarded(apos, 16);
        w[4 * u] = v.x; w[4 * u + 1] = v.y; w[4 * u + 2] = v.z; w[4 * u + 3] = v.w;
      }
      w[16] = load_guarded(span_a + off + 64, 4).x;
    }
  };
  const uint32_t n_blk = span_bytes / S6_BLK;
  constexpr uint32_t BLK_PER_SEG = SEG2_BYTES / S6_BLK;
  load_block(0);
  for (uint32_t blk = 0; blk < n_blk; blk++) {
    if ((blk & (BLK_PER_SEG - 1)) == 0) {
      if (lane == 0) kb[blk / BLK_PER_SEG] = next_k + (q_tail - q_head);
      __syncwarp();
    }
    // flags -> two words, bit 8*byte + word (word 0..7): 32 bytes each.  The cheap zero-byte test can
    // only err on the byte right above a true zero byte, i.e. when two hits touch; that (practically
    // never) redoes the block with the exact test.
    uint32_t e0 = 0, e1 = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      uint32_t z = w[j] ^ pv;
      if (D > 0) z |= __funnelshift_r(w[j], w[j + 1], 8 * D) ^ qv;
      const uint32_t d = (z - 0x01010101u) & ~z & 0x80808080u;
      if (j < 8) e0 |= d >> (7 - j); else e1 |= d >> (15 - j);
    }
    if ((e0 & (e0 >> 8)) | (e1 & (e1 >> 8))) {
      e0 = 0; e1 = 0;
#pragma unroll
      for (int j = 0; j < 16; j++) {
        uint32_t z = w[j] ^ pv;
        if (D > 0) z |= __funnelshift_r(w[j], w[j + 1], 8 * D) ^ qv;
        const uint32_t d = ~(((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z | 0x7F7F7F7Fu);   // 0x80 exactly in the zero bytes
        if (j < 8) e0 |= d >> (7 - j); else e1 |= d >> (15 - j);
      }
    }
    if (blk + 1 < n_blk) load_block(blk + 1);
    if (!interior && (e0 | e1)) {
      // clip to the candidate range [mis, end_a)
      const uint64_t apos = span_a + (uint64_t)blk * S6_BLK + (uint64_t)lane * 64;
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
    const uint32_t any = __ballot_sync(0xFFFFFFFFu, c != 0);
    if (any) {
      const uint32_t pbase0 = blk * S6_BLK + lane * 64;
      uint32_t total;
      if (!__ballot_sync(0xFFFFFFFFu, c > 1)) {
        // at most one hit per lane: the rank among the lanes with a hit is the queue order
        total = __popc(any);
        if (c && q_tail - q_head + total <= S6_QCAP) {
          const uint32_t e = e0 ? e0 : e1;
          const uint32_t b = __ffs(e) - 1;
          Q[(q_tail + __popc(any & lt_mask)) & (S6_QCAP - 1)] = pbase0 + (e0 ? 0u : 32u) + 4 * (b & 7u) + (b >> 3);
        }
      } else {
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += y; }
        total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (c && q_tail - q_head + total <= S6_QCAP) {
          // this lane's hits, inserted in position order (the packed flags are byte-major)
          const uint32_t wb = q_tail + incl - c;
          uint32_t pbase = pbase0;
          uint32_t i = 0, e = e0;
          for (;;) {
            if (!e) { if (!e1) break; e = e1; e1 = 0; pbase += 32; }
            const uint32_t b = __ffs(e) - 1;
            e &= e - 1;
            const uint32_t pos = pbase + 4 * (b & 7u) + (b >> 3);
            uint32_t j = i;
            while (j > 0 && Q[(wb + j - 1) & (S6_QCAP - 1)] > pos) { Q[(wb + j) & (S6_QCAP - 1)] = Q[(wb + j - 1) & (S6_QCAP - 1)]; j--; }
            Q[(wb + j) & (S6_QCAP - 1)] = pos;
            i++;
          }
        }
      }
      if (q_tail - q_head + total > S6_QCAP) dense = true;   // more hits than the ring holds: not this kernel's kind of input
      else q_tail += total;
      __syncwarp();
    }
    const bool last = blk + 1 == n_blk;
    if (last) walk_run(true);
    else if (q_tail - q_head >= 32) walk_run(false);
  }
  if (dense && lane == 0) atomicOr(err, ERR_DENSE);
  if (lane == 0) {
    kb[n_sseg] = next_k;
    for (uint32_t s = 0; s < n_sseg; s++) fb.count[seg0 + s] = min(kb[s + 1] - kb[s], fb.K);
  }
}

}  // namespace rgx
