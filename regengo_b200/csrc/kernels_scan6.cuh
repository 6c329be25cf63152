// findall_scan6_kernel -- the FindAll scan for TDFA patterns with a literal start.  One warp owns a RANGE of up to 32
// consecutive 32 KiB segments; slab entry j of a segment = candidate j: {start_rel, len}, capture offsets.
//
// scan5 (round 1) walked candidates in batches of 32 and was issue-bound with a third of the lanes busy: lanes of
// a batch finish at different times, and every step carried the accept/tag bookkeeping.  Here:
//
//   FILTER    a lane owns 64 contiguous bytes of every 2 KiB block.  Per word, z = (w ^ p) | (w >> 8d ^ q) has a zero
//             byte exactly where byte p is followed, d bytes later, by byte q -- the first and the last of the first
//             four bytes of the literal prefix.  Both are NECESSARY for a match, so no start is lost; the walk checks
//             the rest.  Hits are queued in position order together with their slab index (ballot rank; the general
//             ordered insert only when a lane has two).
//   WALK      a lane is a walker: it takes the next queued candidate the moment its previous walk ends (refill), so
//             walks of different lengths overlap instead of waiting for each other.  Every step is the same code:
//             cell = row[byte] from a register window; the cell carries the shared-memory address of the next row
//             (relocated once per CTA) and an event index.  Index 0 (no tag list fires, nothing about acceptance
//             changes) costs nothing more; any other index appends {index, position} to the candidate's log; a dead
//             cell ends the walk.  What the events MEAN is worked out afterwards.
//   FINALIZE  logs live in a ring of 64 slots indexed by candidate number; as soon as 32 CONSECUTIVE candidates are
//             finished one pass replays their logs -- one candidate per lane, all lanes busy -- and publishes the
//             records.
//   Walker state lives in registers only while walks run; between two walk phases it is parked in shared memory
//   so that the filter has the registers (and the filter's block is not held across a walk phase).
//
// Exactness of reading the match off the log (tdfa.go:929-983).  Between two logged events the walk takes only
// cheap steps; by construction (device_program.cu) those fire no transition list and either stay in the state or
// move between states that never accept.  So after event e the automaton stays in the event's next state S_e for
// as long as it accepts, and the reference re-applies S_e's accept list at every step of that run: only the last
// application, at the position where the next event (or the end of the walk) begins, is observable.  The last
// accepting step of the walk is therefore the end of the run of the LAST event whose next state accepts, and the
// tags at that moment are: for each event up to it, its transition list at its own position, then (if S_e accepts)
// S_e's accept list at the end of its run.
#pragma once
#include "kernels_findall2.cuh"

namespace rgx {

constexpr int S6_SLOTS = 64;             // candidates in flight per warp (power of two)
// LOGCAP (template parameter): events a walk may log, 11 or 15 (the count travels in 4 bits).  A slot is 2 + LOGCAP words:
// queue entry, end word, events (odd: consecutive slots fall into different banks).  The host takes 11 when no walk of
// the pattern can log more (DevMeta::w6_maxev, the longest event path of the automaton), else 15.
constexpr uint32_t S6_QCAP = 128;        // queued candidates per warp (power of two)
constexpr uint32_t S6_BLK = 2048;        // bytes per filter block (64 per lane)
constexpr uint32_t S6_MAXR = 32;         // segments per range (queue entries hold a 20-bit range-relative start)
constexpr uint32_t S6_MAXJ = 4096;       // slab entries per segment the queue entry can number
constexpr int S6_SAVE = 8;               // parked walker state, words per lane

__host__ __device__ inline size_t scan6_warp_words(int ntags, int logcap) {
  return (size_t)S6_SLOTS * (2 + logcap) + S6_QCAP + (size_t)(ntags > S6_SAVE ? ntags : S6_SAVE) * 32;
}

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// PLEN: prefix bytes the filter looks at (1..4; of four it tests bytes 0, 1 and 3); 0: a first-byte set of one or two
// ASCII ranges instead of a literal.  WARPS warps per CTA, GROUPS
// window words per walk iteration.
template <int PLEN, int WARPS, int MINB, int GROUPS, bool PF, int LOGCAP>
__global__ void __launch_bounds__(WARPS * 32, MINB) findall_scan6_kernel(
    const DevMeta m, const uint32_t* __restrict__ gimg, const uint8_t* __restrict__ buf, const uint64_t len,
    const uint32_t mis, const uint64_t n_seg, const uint32_t R, const uint32_t walk_at, const FindAllBufs fb, int* err) {
  constexpr int S6_LOGCAP = LOGCAP, S6_SLOTW = 2 + LOGCAP;
  extern __shared__ __align__(16) uint32_t smem_all[];
  __shared__ __align__(8) unsigned long long mbar;
  stage_image_tma(smem_all, gimg + m.w6_off, m.w6_words, &mbar);
  uint32_t rows_s = smem_u32(smem_all);
  asm volatile("" : "+r"(rows_s));   // (opaque: the compiler would otherwise re-derive the shared-window address at every use)
  // relocation: the row field of every live cell becomes an absolute shared-memory address
  {
    const uint32_t n_cells = (uint32_t)m.t_ns * 256u;
    for (uint32_t i = threadIdx.x; i < n_cells; i += WARPS * 32) {
      const uint32_t c = smem_all[i];
      if (c < S6_DEAD) smem_all[i] = c + rows_s;
    }
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int nt = m.t_ntags;
  uint32_t* wbase = smem_all + m.w6_words + (size_t)warp * scan6_warp_words(nt, LOGCAP);
  uint32_t slots_s = smem_u32(wbase);                                     // S6_SLOTS x S6_SLOTW words
  asm volatile("" : "+r"(slots_s));
  uint32_t* Q = wbase + S6_SLOTS * S6_SLOTW;                               // S6_QCAP entries: j << 20 | range-relative start
  uint32_t* SV = Q + S6_QCAP;                                              // parked walker state: SV[i * 32 + lane]
  int32_t* T = reinterpret_cast<int32_t*>(SV) + lane;                      // tags of the replay (same words): T[j * 32]
  uint32_t desc_s = rows_s + m.w6_desc * 4u;                               // descriptor d at desc_s + 32 d: flags, then its tag entries
  uint32_t tags_s = slots_s + (S6_SLOTS * S6_SLOTW + S6_QCAP) * 4u + lane * 4u;   // this lane's tag j is at tags_s + 128 j
  asm volatile("" : "+r"(desc_s), "+r"(tags_s));
  const uint32_t* initt = smem_all + m.w6_init;
  const uint32_t start_row = rows_s + (uint32_t)m.t_start_any * 1024u;
  const uint8_t* abuf = buf - mis;                        // 16-byte aligned view of the buffer
  const uint64_t end_a = (uint64_t)mis + fb.cand_len;     // candidate starts are in [mis, end_a)
  const uint64_t load_end = (uint64_t)mis + len;          // bytes exist in [mis, load_end) (shard + halo)
  const uint32_t pv0 = (uint32_t)m.prefix_bytes[0] * 0x01010101u, pv1 = (uint32_t)m.prefix_bytes[1] * 0x01010101u;
  const uint32_t pv2 = (uint32_t)m.prefix_bytes[PLEN == 4 ? 3 : 2] * 0x01010101u;
  // range filter (PLEN == 0): x + (0x80 - lo) has bit 7 iff x >= lo, x + (0x7F - hi) iff x > hi (x <= 0x7F)
  const uint32_t ge0 = (0x80u - m.w6_rlo[0]) * 0x01010101u, gt0 = (0x7Fu - m.w6_rhi[0]) * 0x01010101u;
  const uint32_t ge1 = (0x80u - m.w6_rlo[1]) * 0x01010101u, gt1 = (0x7Fu - m.w6_rhi[1]) * 0x01010101u;
  const bool two_ranges = m.w6_nrng > 1;
  const uint64_t n_ranges = (n_seg + R - 1) / R;
  const uint64_t total_warps = (uint64_t)gridDim.x * WARPS;

  for (uint64_t range = (uint64_t)blockIdx.x * WARPS + warp; range < n_ranges; range += total_warps) {
    const uint64_t seg0 = range * R;
    const uint32_t n_rseg = (uint32_t)min((uint64_t)R, n_seg - seg0);
    const uint64_t base_a = seg0 * SEG2_BYTES;
    const uint8_t* sp = abuf + base_a;                    // range base; every position below is relative to it
    asm volatile("" : "+l"(sp));                          // (opaque to the compiler: keep it in registers, do not re-derive it)
    const uint32_t range_bytes = n_rseg * SEG2_BYTES;
    const uint64_t avail64 = load_end - base_a;
    const uint32_t lim = avail64 > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t)avail64;   // bytes readable from sp
    const uint32_t lim_eot = avail64 <= 0xFFFFFFF0ull ? lim : 0xFFFFFFFFu;            // position that means "the buffer's last byte was consumed"

    // ---- warp-uniform bookkeeping ----
    uint32_t q_head = 0, q_tail = 0;      // queue ring counters
    uint32_t next_k = 0, fin_base = 0;    // next candidate number to hand out; first candidate not yet finalized
    uint32_t done_lo = 0, done_hi = 0;    // finished flags of candidates fin_base + 0..31 / + 32..63
    uint32_t seg_count = 0;               // candidates of the current segment so far
    bool dense = false;
    SV[7 * 32 + lane] = 0;                // every walker idle
    __syncwarp();

    // FINALIZE candidates [base, base + cnt): replay the logs, publish the records
    auto finalize = [&](const uint32_t base, const uint32_t cnt) {
      const uint32_t kk = base + lane;
      const bool valid = (uint32_t)lane < cnt;
      const uint32_t ss = slots_s + (kk & (S6_SLOTS - 1)) * (S6_SLOTW * 4u);
      const uint32_t qe = valid ? lds32(ss) : 0u;
      const uint32_t hdr = valid ? lds32(ss + 4) : 0u;
      const uint32_t st = qe & 0xFFFFFu, j = qe >> 20;
      // (log overflow or a walk of a MiB: the generic scan decides, and nothing of that log is read)
      const uint32_t nl = (hdr >> 31) ? 0u : ((hdr >> 22) & 15u), end_rel = (hdr & 0x3FFFFFu) - st;
      if (hdr >> 31) atomicOr(err, ERR_DENSE);
      // the last event whose next state accepts (acceptStatesEOT counts when the walk consumed the input's last byte)
      uint32_t la = 0;
      uint32_t nmax = __reduce_max_sync(0xFFFFFFFFu, nl);
      for (uint32_t e = 0; e < nmax; e++)
        if (e < nl && (lds32(desc_s + (lds32(ss + 8 + e * 4) >> 22) * 32u) & S6_ACC)) la = e + 1;
      if (((hdr >> 30) & 1u) && nl && (lds32(desc_s + (lds32(ss + 8 + (nl - 1) * 4) >> 22) * 32u) & S6_ACC_EOT)) la = nl;
      const bool matched = valid && la > 0;
      const int32_t match_end = !matched ? -1 : (la < nl ? (int32_t)((lds32(ss + 8 + la * 4) & 0x3FFFFFu) - st) : (int32_t)end_rel);
      for (int q = 0; q < nt; q++) T[q * 32] = -1;
      T[0] = 0;
      for (int t = 0; t < m.t_n_init_any; t++) T[initt[t] * 32] = 0;
      // replay events 0 .. la-1: transition actions at the step's position, accept actions at the end of the run
      nmax = __reduce_max_sync(0xFFFFFFFFu, la);
      uint32_t ev = la ? lds32(ss + 8) : 0u;
      for (uint32_t e = 0; e < nmax; e++) {
        if (e < la) {
          const uint32_t idx = ev >> 22;
          const int32_t pos = (int32_t)((ev & 0x3FFFFFu) - st) + 1;
          int32_t run_end = (int32_t)end_rel;
          if (e + 1 < nl) { ev = lds32(ss + 12 + e * 4); run_end = (int32_t)((ev & 0x3FFFFFu) - st); }
          const uint32_t da = desc_s + idx * 32u;
          const uint32_t d = lds32(da);
          const bool accp = (d & S6_ACC) || e + 1 == la;
          const uint32_t n = d >> 24;
          for (uint32_t i = 0; i < n; i++) {
            const uint32_t en = lds32(da + 4u + i * 4u);
            const bool acc_ent = (en & S6_ENT_ACCEPT) != 0;
            if (!acc_ent || accp) sts32(tags_s + (en & 0xFFFFu), (uint32_t)((acc_ent ? run_end : pos) - (int32_t)((en >> 16) & 0xFFu)));
          }
        }
      }
      if (valid && j < fb.K) {
        const uint64_t r = (seg0 + (st >> 15)) * fb.K + j;
        if (matched) {
          fb.keys[r] = make_uint2(st & (SEG2_BYTES - 1), (uint32_t)match_end);
          int32_t* cp = fb.caps + r * fb.cw;
          for (int q = 2; q + 3 < nt && !(fb.cw & 3u); q += 4) {   // two groups per 16-byte store
            int4 v = make_int4(T[q * 32], T[(q + 1) * 32], T[(q + 2) * 32], T[(q + 3) * 32]);
            if (v.x >= 0 && v.y < 0) v.y = match_end;  // unset group end := match end (tdfa.go:1039-1041)
            if (v.z >= 0 && v.w < 0) v.w = match_end;
            *reinterpret_cast<int4*>(cp + (q - 2)) = v;
          }
          for (int q = (fb.cw & 3u) ? 2 : 2 + ((nt - 2) & ~3); q < nt; q += 2) {
            const int32_t a = T[q * 32];
            int32_t b = T[(q + 1) * 32];
            if (a >= 0 && b < 0) b = match_end;
            cp[q - 2] = a;
            cp[q - 1] = b;
          }
        } else {
          fb.keys[r] = make_uint2(st & (SEG2_BYTES - 1), KEY_INVALID);
        }
      }
      __syncwarp();
    };

    // WALK phase: iterate while candidates are queued (drain: until every walk has ended and every record is out)
    auto walk_run = [&](const bool drain) {
      // walker state back into registers
      uint32_t ri = SV[0 * 32 + lane], row = SV[1 * 32 + lane], lp = SV[2 * 32 + lane], lend = SV[3 * 32 + lane];
      uint32_t x0 = SV[4 * 32 + lane], x1 = SV[5 * 32 + lane], x2 = SV[6 * 32 + lane], fl = SV[7 * 32 + lane];   // fl: 1 active, 2 slow (no window)
      __syncwarp();
      bool tail = false;                  // drain: every walk has ended, only partial groups are left to publish
      for (;;) {
        const uint32_t idle_m = __ballot_sync(0xFFFFFFFFu, !(fl & 1u));
        const uint32_t qn = q_tail - q_head;
        if (qn == 0) {
          if (!drain) break;
          tail = idle_m == 0xFFFFFFFFu;
        } else if (idle_m) {
          // refill: the idle lanes take the next candidates, in lane order (a slot must be free: fin_base + 64 > k)
          const uint32_t navail = min(qn, fin_base + S6_SLOTS - next_k);
          const uint32_t rank = __popc(idle_m & lt_mask);
          if (!(fl & 1u) && rank < navail) {
            const uint32_t qe = Q[(q_head + rank) & (S6_QCAP - 1)];
            const uint32_t ss = slots_s + ((next_k + rank) & (S6_SLOTS - 1)) * (S6_SLOTW * 4u);
            sts32(ss, qe);
            ri = qe & 0xFFFFFu; row = start_row; lp = ss + 8; lend = ss + S6_SLOTW * 4u; fl = 1;
            const uint32_t a = ri & ~3u;
            if (a + 12 <= lim) {
              const uint32_t* wp = reinterpret_cast<const uint32_t*>(sp + a);
              x0 = wp[0]; x1 = wp[1]; x2 = wp[2];
            } else fl = 3;
          }
          const uint32_t taken = min((uint32_t)__popc(idle_m), navail);
          q_head += taken; next_k += taken;
        }
        // steps
        bool alive = (fl & 1u) != 0, eob = false;
        auto step = [&](const uint32_t b4) {   // b4 = 4 * byte
          const uint32_t cell = lds32(row + b4);
          if (cell >= S6_DEAD) { alive = false; return; }
          // an event: {index, position} into the log (predicated, no branch); lp keeps counting past the end
          asm volatile(
              "{\n\t.reg .pred p, q;\n\t"
              "setp.ge.u32 p, %1, %4;\n\t"
              "setp.lt.and.u32 q, %0, %2, p;\n\t"
              "@q st.shared.u32 [%0], %3;\n\t"
              "@p add.u32 %0, %0, 4;\n\t}"
              : "+r"(lp) : "r"(cell), "r"(lend), "r"((cell & 0xFFC00000u) | ri), "n"(S6_EVMIN) : "memory");
          row = cell & 0x3FFFFFu;
          ri++;
        };
#pragma unroll
        for (int g = 0; g < GROUPS; g++) {
          if (alive) {
            if (!(fl & 2u)) {
              const uint32_t win = __funnelshift_r(x0, x1, (ri & 3u) * 8u);
              const uint32_t a = ri & ~3u;
              step((win << 2) & 0x3FCu);
              if (alive) step((win >> 6) & 0x3FCu);
              if (alive) step((win >> 14) & 0x3FCu);
              if (alive) step((win >> 22) & 0x3FCu);
              if (alive) {
                x0 = x1; x1 = x2;
                if (a + 16 <= lim) x2 = *reinterpret_cast<const uint32_t*>(sp + a + 12); else fl |= 2u;
              }
            } else {
              for (int s = 0; s < 4 && alive; s++) {
                if (ri >= lim) { alive = false; eob = true; }
                else step((uint32_t)sp[ri] * 4u);
              }
            }
          }
        }
        // the walks that ended: one word says where and with how many events
        const bool fin_now = (fl & 1u) && !alive;
        if (fin_now) {
          const uint32_t nlog = (lp - (lend - S6_LOGCAP * 4u)) >> 2;
          uint32_t hdr = (ri & 0x3FFFFFu) | (min(nlog, 15u) << 22);
          if (eob && ri == lim_eot) hdr |= 1u << 30;
          if (ri >= (1u << 22) || nlog > (uint32_t)S6_LOGCAP) hdr |= 1u << 31;
          if (eob && fb.not_last) atomicOr(err, ERR_HALO);   // ran off the halo: this shard cannot decide the match alone
          sts32(lend - S6_SLOTW * 4u + 4, hdr);
          fl = 0;
        }
        const uint32_t fm = __ballot_sync(0xFFFFFFFFu, fin_now);
        if (fm) {
          // candidate number (mod 64) from the slot address
          const uint32_t b = ((lend - slots_s) / (S6_SLOTW * 4u) - 1u - fin_base) & (S6_SLOTS - 1);
          done_lo |= __reduce_or_sync(0xFFFFFFFFu, (fin_now && b < 32) ? (1u << b) : 0u);
          done_hi |= __reduce_or_sync(0xFFFFFFFFu, (fin_now && b >= 32) ? (1u << (b - 32)) : 0u);
        }
        // (the replay's tag words share the parking area: it is free while walks run)
        if (done_lo == 0xFFFFFFFFu || tail) __syncwarp();   // the logs and headers other lanes wrote are read below
        while (done_lo == 0xFFFFFFFFu || (tail && fin_base < next_k)) {
          finalize(fin_base, min(32u, next_k - fin_base));
          fin_base += 32; done_lo = done_hi; done_hi = 0;
        }
        if (tail) break;
      }
      if (!drain) {
        SV[0 * 32 + lane] = ri; SV[1 * 32 + lane] = row; SV[2 * 32 + lane] = lp; SV[3 * 32 + lane] = lend;
        SV[4 * 32 + lane] = x0; SV[5 * 32 + lane] = x1; SV[6 * 32 + lane] = x2; SV[7 * 32 + lane] = fl;
      }
      __syncwarp();
    };

    // ---------------- FILTER ----------------
    // A block that lies wholly inside the candidate range (and whose look-ahead word exists) takes the unrolled
    // path; the block at either end of the buffer is tested byte by byte with bounds checks (at most two per call).
    // A prefix cut off by the end of the buffer cannot match: every prefix state is non-accepting.
    const uint32_t n_blk = range_bytes / S6_BLK;
    constexpr uint32_t BLK_PER_SEG = SEG2_BYTES / S6_BLK;
    for (uint32_t blk = 0; blk < n_blk; blk++) {
      const uint32_t off = blk * S6_BLK + lane * 64;
      const uint64_t blk_a = base_a + (uint64_t)blk * S6_BLK;
      uint32_t e0 = 0, e1 = 0;   // flags, bit 8*byte + word (word 0..7): 32 bytes each
      if (blk_a >= mis && blk_a + S6_BLK <= end_a && blk_a + S6_BLK + 4 <= load_end) {
        uint32_t w[17];
        // the block after this one into L2 -> L1 while this one is filtered and walked (no registers held across
        // the walk phase: a prefetch, not a load)
        if (PF && blk_a + 2 * S6_BLK <= load_end) asm volatile("prefetch.global.L1 [%0];" ::"l"(sp + off + S6_BLK));
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const uint4 v = *reinterpret_cast<const uint4*>(sp + off + u * 16);
          w[4 * u] = v.x; w[4 * u + 1] = v.y; w[4 * u + 2] = v.z; w[4 * u + 3] = v.w;
        }
        w[16] = PLEN > 1 ? *reinterpret_cast<const uint32_t*>(sp + off + 64) : 0u;
        // The cheap zero-byte test can only err on the byte right above a true zero byte, i.e. when two hits
        // touch; that (practically never) redoes the block with the exact test.
#pragma unroll
        for (int q = 0; q < 16; q++) {
          uint32_t d;
          if (PLEN == 0) {
            // first-byte SET: 0x80 in every byte of the word that lies in [lo, hi] (exact: no carries between bytes)
            const uint32_t x = w[q] & 0x7F7F7F7Fu;
            d = (x + ge0) & ~(x + gt0);
            if (two_ranges) d |= (x + ge1) & ~(x + gt1);
            d &= ~w[q] & 0x80808080u;
          } else {
            uint32_t z = w[q] ^ pv0;
            if (PLEN > 1) z |= __funnelshift_r(w[q], w[q + 1], 8) ^ pv1;
            if (PLEN > 2) z |= __funnelshift_r(w[q], w[q + 1], PLEN == 4 ? 24 : 16) ^ pv2;
            d = (z - 0x01010101u) & ~z & 0x80808080u;
          }
          if (q < 8) e0 |= d >> (7 - q); else e1 |= d >> (15 - q);
        }
        if (PLEN > 0 && ((e0 & (e0 >> 8)) | (e1 & (e1 >> 8)))) {
          e0 = 0; e1 = 0;
#pragma unroll
          for (int q = 0; q < 16; q++) {
            uint32_t z = w[q] ^ pv0;
            if (PLEN > 1) z |= __funnelshift_r(w[q], w[q + 1], 8) ^ pv1;
            if (PLEN > 2) z |= __funnelshift_r(w[q], w[q + 1], PLEN == 4 ? 24 : 16) ^ pv2;
            const uint32_t d = ~(((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z | 0x7F7F7F7Fu);   // 0x80 exactly in the zero bytes
            if (q < 8) e0 |= d >> (7 - q); else e1 |= d >> (15 - q);
          }
        }
      } else if (blk_a < end_a) {
        for (uint32_t i = 0; i < 64; i++) {
          const uint64_t ap = blk_a + (uint64_t)lane * 64 + i;
          bool hit = ap >= mis && ap < end_a;
          if (hit && PLEN == 0) { const uint32_t c = abuf[ap]; hit = (c >= m.w6_rlo[0] && c <= m.w6_rhi[0]) || (two_ranges && c >= m.w6_rlo[1] && c <= m.w6_rhi[1]); }
          if (hit && PLEN > 0) hit = abuf[ap] == m.prefix_bytes[0];
          if (PLEN > 1) hit = hit && ap + 1 < load_end && abuf[ap + 1] == m.prefix_bytes[1];
          if (PLEN > 2) hit = hit && ap + (PLEN == 4 ? 3 : 2) < load_end && abuf[ap + (PLEN == 4 ? 3 : 2)] == m.prefix_bytes[PLEN == 4 ? 3 : 2];
          if (hit) { if (i < 32) e0 |= 1u << (8 * (i & 3) + (i >> 2)); else e1 |= 1u << (8 * (i & 3) + ((i - 32) >> 2)); }
        }
      }
      if ((blk & (BLK_PER_SEG - 1)) == 0 && blk) {
        if (lane == 0) fb.count[seg0 + blk / BLK_PER_SEG - 1] = min(seg_count, fb.K);
        seg_count = 0;
      }
      const uint32_t c = __popc(e0) + __popc(e1);
      const uint32_t any = __ballot_sync(0xFFFFFFFFu, c != 0);
      if (any) {
        uint32_t total;
        if (!__ballot_sync(0xFFFFFFFFu, c > 1)) {
          // at most one hit per lane: the rank among the lanes with a hit is the queue order
          total = __popc(any);
          if (c && q_tail - q_head + total <= S6_QCAP) {
            const uint32_t e = e0 ? e0 : e1;
            const uint32_t b = __ffs(e) - 1;
            const uint32_t rank = __popc(any & lt_mask);
            Q[(q_tail + rank) & (S6_QCAP - 1)] = ((seg_count + rank) << 20) | (off + (e0 ? 0u : 32u) + 4 * (b & 7u) + (b >> 3));
          }
        } else {
          // hits before this lane: ballots of the bit planes of c (a lane with 8 or more takes the shuffle scan)
          uint32_t incl;
          if (!__ballot_sync(0xFFFFFFFFu, c > 7)) {
            const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, c & 1u), b1 = __ballot_sync(0xFFFFFFFFu, c & 2u), b2 = __ballot_sync(0xFFFFFFFFu, c & 4u);
            incl = c + __popc(b0 & lt_mask) + 2 * __popc(b1 & lt_mask) + 4 * __popc(b2 & lt_mask);
            total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
          } else {
            incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += y; }
            total = __shfl_sync(0xFFFFFFFFu, incl, 31);
          }
          if (c && q_tail - q_head + total <= S6_QCAP) {
            // this lane's hits in position order (the packed flags are byte-major: bit 8*byte + word): lowest word
            // with a flag first, then that word's bytes
            uint32_t i = incl - c;
            for (uint32_t half = 0; half < 2; half++) {
              uint32_t e = half ? e1 : e0;
              while (e) {
                const uint32_t wd = __ffs((e | (e >> 8) | (e >> 16) | (e >> 24)) & 0xFFu) - 1;
                uint32_t nib = (e >> wd) & 0x01010101u;
                e &= ~(0x01010101u << wd);
                while (nib) {
                  const uint32_t by = (__ffs(nib) - 1) >> 3;
                  nib &= nib - 1;
                  Q[(q_tail + i) & (S6_QCAP - 1)] = ((seg_count + i) << 20) | (off + half * 32 + 4 * wd + by);
                  i++;
                }
              }
            }
          }
        }
        if (q_tail - q_head + total > S6_QCAP) dense = true;   // more hits than the ring holds: not this kernel's kind of input
        else q_tail += total;
        seg_count += total;
        if (seg_count > min(fb.K, S6_MAXJ) && lane == 0) atomicOr(err, seg_count > S6_MAXJ ? ERR_DENSE : ERR_SLAB);   // records beyond the slab are dropped
        __syncwarp();
      }
      const bool last = blk + 1 == n_blk;
      if (last || q_tail - q_head >= walk_at) walk_run(last);
    }
    if (lane == 0) {
      fb.count[seg0 + n_rseg - 1] = min(seg_count, fb.K);
      if (dense) atomicOr(err, ERR_DENSE);
    }
    __syncwarp();
  }
}

}  // namespace rgx
