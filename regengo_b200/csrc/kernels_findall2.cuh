// FindAllBytes, second-generation kernels.
//
//  findall_scan_tdfa_kernel   the HBM-bound scan for TDFA patterns whose start filter is a literal
//                             prefix of >= 2 bytes (e.g. "ht" of https?://...).  One warp per 32 KiB
//                             segment.  FILTER: lanes stream the segment with 16-byte coalesced loads,
//                             four in flight per lane, and run a two-byte SIMD-in-register compare
//                             (3 ALU ops per word and pattern byte); hits are queued in order with a
//                             warp prefix sum.  WALK: the queued candidates are verified 32 at a time,
//                             one TDFA walk per lane, each step reading one fused shared-memory cell
//                             (next state, transition tag list, accept flag and accept tag list of the
//                             next state).  Tags live in shared memory ([tag][lane], conflict free);
//                             the accept snapshot is copy-on-write and accept tag lists are applied
//                             lazily (only when the list or a transition list changes), so the common
//                             step is a byte load, a cell load and a handful of ALU ops.
//  findall_chain2_kernel      cursor replay over the record slabs (see kernels_findall.cuh), one lane
//                             per part, 32-bit division fast path, per-segment output bases.
//  findall_emit2_kernel       one warp per segment writes the kept records in order.
//
// Slab entry j of a segment is CANDIDATE j (key.y == KEY_INVALID when it did not match), so no
// in-order compaction is needed in the scan; the chain/emit kernels skip invalid entries.
#pragma once
#include "engines.cuh"
#include "kernels_findall.cuh"

namespace rgx {

constexpr uint32_t SEG2_BYTES = 32768;    // bytes per segment (one warp) in the v2 scan
constexpr uint32_t Q2CAP = 2048;          // per-warp candidate queue (u16 segment-relative positions)
constexpr uint32_t KEY_INVALID = 0xFFFFFFFFu;
constexpr int SCAN2_WARPS = 8;
constexpr int NT_MAX2 = 16;               // tags handled by the fast walk
constexpr int FILTER_UNROLL = 4;

// approximate per-byte equality: bit 7 of every byte of x that equals pat's byte is set; a byte
// above a matching byte may be flagged too (false positives only -- candidates are verified)
__device__ __forceinline__ uint32_t eq_approx(uint32_t w, uint32_t pat) {
  const uint32_t x = w ^ pat;
  return (x - 0x01010101u) & ~x & 0x80808080u;
}

// dynamic shared memory: [program image][tags: warps x ntags x 32 ints][event logs: warps x LOG2CAP x 32 x 8 B]
constexpr int LOG2CAP = 12;
__host__ __device__ inline size_t scan2_tag_words(int ntags) {
  size_t w = (size_t)SCAN2_WARPS * ntags * 32;
  w = (w + 1) & ~(size_t)1;   // keep the uint2 log 8-byte aligned
  return w + (size_t)SCAN2_WARPS * LOG2CAP * 32 * 2;
}

__global__ void __launch_bounds__(SCAN2_WARPS * 32, 3) findall_scan_tdfa_kernel(
    const DevMeta m, const uint32_t* __restrict__ gimg, const uint8_t* __restrict__ buf, const uint64_t len,
    const uint32_t mis, const uint64_t n_seg, const FindAllBufs fb, int* err) {
  extern __shared__ __align__(16) uint32_t smem_all[];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint16_t queue[SCAN2_WARPS][Q2CAP];
  stage_image_tma(smem_all, gimg, m.image_words, &mbar);
  const uint32_t* img = smem_all;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nt = m.t_ntags;
  int32_t* T = reinterpret_cast<int32_t*>(smem_all + m.image_words) + (size_t)warp * nt * 32 + lane;      // tags: T[j*32]
  uint2* LG = reinterpret_cast<uint2*>(reinterpret_cast<int32_t*>(smem_all + m.image_words) + (size_t)SCAN2_WARPS * nt * 32) +
              (size_t)warp * LOG2CAP * 32 + lane;                                                         // event log: LG[e*32]
  uint16_t* q = queue[warp];
  const uint8_t* abuf = buf - mis;
  const uint64_t end_a = (uint64_t)mis + fb.cand_len;    // candidate starts are in [mis, end_a)
  const uint64_t load_end = (uint64_t)mis + len;         // bytes exist in [mis, load_end) (shard + halo)
  const uint32_t p0 = (uint32_t)m.prefix_bytes[0] * 0x01010101u;
  const uint32_t p1 = (uint32_t)m.prefix_bytes[1] * 0x01010101u;
  const uint32_t* fast = img + m.off_t_fast;
  const uint32_t* aoff = img + m.off_t_alist_off;
  const uint32_t* alist = img + m.off_t_alist;
  const uint64_t total_warps = (uint64_t)gridDim.x * SCAN2_WARPS;
  const int64_t l = (int64_t)len;
  constexpr uint32_t N_IT = SEG2_BYTES / 512;

  for (uint64_t seg = (uint64_t)blockIdx.x * SCAN2_WARPS + warp; seg < n_seg; seg += total_warps) {
    const uint64_t seg_a = seg * SEG2_BYTES;
    const bool interior = seg_a >= mis && seg_a + SEG2_BYTES <= end_a;
    uint32_t tail = 0, slot_base = 0;
    for (uint32_t it0 = 0; it0 <= N_IT; it0 += FILTER_UNROLL) {
      if (it0 < N_IT) {
        uint4 v[FILTER_UNROLL];
        if (interior) {
#pragma unroll
          for (int u = 0; u < FILTER_UNROLL; u++)
            v[u] = *reinterpret_cast<const uint4*>(abuf + seg_a + (uint64_t)(it0 + u) * 512 + (uint64_t)lane * 16);
        } else {
#pragma unroll
          for (int u = 0; u < FILTER_UNROLL; u++) {
            const uint64_t apos = seg_a + (uint64_t)(it0 + u) * 512 + (uint64_t)lane * 16;
            v[u] = make_uint4(0, 0, 0, 0);
            if (apos >= mis && apos + 16 <= load_end) {
              v[u] = *reinterpret_cast<const uint4*>(abuf + apos);
            } else if (apos + 16 > mis && apos < load_end) {
              uint8_t* vb = reinterpret_cast<uint8_t*>(&v[u]);
              for (int j = 0; j < 16; j++) if (apos + j >= mis && apos + j < load_end) vb[j] = abuf[apos + j];
            }
          }
        }
#pragma unroll
        for (int u = 0; u < FILTER_UNROLL; u++) {
          const uint32_t t0 = eq_approx(v[u].x, p1), t1 = eq_approx(v[u].y, p1), t2 = eq_approx(v[u].z, p1), t3 = eq_approx(v[u].w, p1);
          // second-byte flags of the NEXT lane's first word; the last lane cannot see its successor
          // and keeps its 16th byte as a candidate on the first byte alone
          uint32_t t4 = __shfl_down_sync(0xFFFFFFFFu, t0, 1);
          if (lane == 31) t4 = 0x80u;
          const uint32_t c0 = eq_approx(v[u].x, p0) & __funnelshift_r(t0, t1, 8);
          const uint32_t c1 = eq_approx(v[u].y, p0) & __funnelshift_r(t1, t2, 8);
          const uint32_t c2 = eq_approx(v[u].z, p0) & __funnelshift_r(t2, t3, 8);
          const uint32_t c3 = eq_approx(v[u].w, p0) & __funnelshift_r(t3, t4, 8);
          const bool any = (c0 | c1 | c2 | c3) != 0;
          const uint32_t bal = __ballot_sync(0xFFFFFFFFu, any);
          if (bal) {
            const uint32_t it = it0 + u;
            uint32_t mask = 0;
            if (any) {
              mask = gather4(c0) | (gather4(c1) << 4) | (gather4(c2) << 8) | (gather4(c3) << 12);
              if (!interior) {
                const uint64_t apos = seg_a + (uint64_t)it * 512 + (uint64_t)lane * 16;
                if (apos >= end_a) mask = 0;
                else {
                  if (apos < mis) mask = (mis - apos >= 16) ? 0u : (mask & ~((1u << (uint32_t)(mis - apos)) - 1u));
                  if (apos + 16 > end_a) mask &= (1u << (uint32_t)(end_a - apos)) - 1u;
                }
              }
            }
            const uint32_t c = __popc(mask);
            uint32_t incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += y; }
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            uint32_t w = tail + incl - c;
            if (tail + total <= Q2CAP) {
              while (mask) {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1;
                q[w++] = (uint16_t)(it * 512 + lane * 16 + j);
              }
              tail += total;
            } else if (lane == 0) {
              atomicOr(err, ERR_DENSE);  // more than Q2CAP candidates in 32 KiB: the host re-runs the generic scan
            }
          }
        }
        __syncwarp();
      }
      // verify the queued candidates at the end of the segment
      if (it0 >= N_IT && tail > 0) {
        const uint32_t n = tail;
        for (uint32_t base = 0; base < n; base += 32) {
          const uint32_t k = base + lane;
          bool active = k < n;
          const uint32_t srel = active ? q[k] : 0;
          const int64_t s = (int64_t)(seg_a + srel) - (int64_t)mis;
          // WALK: state transitions only.  Tag lists that fire are appended to a per-lane event log
          // {list id, position}; accept lists are logged lazily (when the list changes).  The tags are
          // rebuilt from the log afterwards, with all lanes converged.
          int64_t i = s;
          uint32_t state = (uint32_t)m.t_start_any;
          int32_t match_end = -1;   // relative to s
          uint32_t pend_al = 0;     // accept tag list in force at match_end
          uint32_t nlog = 0;
          uint32_t word = 0;        // the aligned 4 input bytes that contain byte i
          bool have = false;
          while (__any_sync(0xFFFFFFFFu, active)) {
            if (active) {
              uint32_t cell = FAST_NONE;
              if (i < l) {
                if (!have || ((i + mis) & 3) == 0) { word = *reinterpret_cast<const uint32_t*>(abuf + (((uint64_t)(i + mis)) & ~3ull)); have = true; }
                const uint32_t c = (word >> ((uint32_t)((i + mis) & 3) * 8)) & 255u;
                if (c < 128) cell = fast[state * 128 + c];
              } else if (fb.not_last) {
                atomicOr(err, ERR_HALO);  // a walk ran off the halo: the shard cannot decide this match alone
              }
              if ((cell & 0x3FFu) == FAST_NONE) {
                active = false;
              } else {
                const int32_t pos = (int32_t)(i + 1 - s);  // tag value base: i + 1 - start
                const uint32_t al = (cell >> 10) & 0x3FFu;
                if (al) {
                  if (pend_al) { if (nlog < LOG2CAP) LG[nlog * 32] = make_uint2(pend_al, (uint32_t)match_end); nlog++; pend_al = 0; }
                  if (nlog < LOG2CAP) LG[nlog * 32] = make_uint2(al, (uint32_t)pos);
                  nlog++;
                }
                state = cell & 0x3FFu;
                if ((cell >> 30) && ((cell & (1u << 30)) || i == l - 1)) {
                  const uint32_t aal = (cell >> 20) & 0x3FFu;
                  if (aal != pend_al) {
                    if (pend_al) { if (nlog < LOG2CAP) LG[nlog * 32] = make_uint2(pend_al, (uint32_t)match_end); nlog++; }
                    pend_al = aal;
                  }
                  match_end = pos;
                }
                i++;
              }
            }
          }
          if (nlog > LOG2CAP) atomicOr(err, ERR_DENSE);   // more tag events than the log holds: generic scan instead
          // REPLAY: tags := -1; apply, in order, every logged list whose position is <= match_end (events
          // after the last accept never reached the snapshot, tdfa.go:963-975), then the accept list in force.
          for (int j = 0; j < nt; j++) T[j * 32] = -1;
          T[0] = 0;
          for (int t = 0; t < m.t_n_init_any; t++) T[img[m.off_t_init + m.t_n_init_begin + t] * 32] = 0;
          const bool matched = k < n && match_end >= 0;
          uint32_t nmax = matched ? min(nlog, (uint32_t)LOG2CAP) : 0;
#pragma unroll
          for (int o = 16; o; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xFFFFFFFFu, nmax, o));
          const uint32_t nmine = matched ? min(nlog, (uint32_t)LOG2CAP) : 0;
          for (uint32_t e = 0; e < nmax; e++) {
            if (e < nmine) {
              const uint2 ev = LG[e * 32];
              if ((int32_t)ev.y <= match_end)
                for (uint32_t a = aoff[ev.x]; a < aoff[ev.x + 1]; a++) { const uint32_t x = alist[a]; T[(x & 0xFFFFu) * 32] = (int32_t)ev.y - (int32_t)(x >> 16); }
            }
          }
          if (matched && pend_al)
            for (uint32_t a = aoff[pend_al]; a < aoff[pend_al + 1]; a++) { const uint32_t x = alist[a]; T[(x & 0xFFFFu) * 32] = match_end - (int32_t)(x >> 16); }
          // publish candidate k
          if (k < n) {
            const uint64_t r = seg * fb.K + slot_base + k;
            if (slot_base + k < fb.K) {
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
        slot_base += n;
        tail = 0;
      }
    }
    if (lane == 0) fb.count[seg] = min(slot_base, fb.K);
  }
}

// ---- chain over slabs (valid or invalid entries), one lane per part ---------------------------------
struct Chain2Bufs {
  long long* exit_prev;
  long long* exit_cur;
  unsigned long long* part_sel;
  unsigned long long* part_reps;
  uint32_t* seg_sel;              // [n_seg] kept records before this segment, inside its part
  unsigned long long* seg_reps;   // [n_seg] matches returned before this segment, inside its part
  int* changed;                   // [pass] set when a part's exit differs from the previous pass
};

// One lane per part, one warp per 32 consecutive parts.  The lanes' record streams are staged through
// shared memory in chunks of CH2 records with COALESCED cooperative loads (a lane reading its own slab
// directly would take a DRAM miss on almost every step); each lane then replays its chunk from
// shared memory.
constexpr int CH2 = 32;
constexpr int CHAIN2_WARPS = 2;

template <int ENGINE>
__global__ void __launch_bounds__(CHAIN2_WARPS * 32) findall_chain2_kernel(
    const uint64_t n_seg, const uint32_t seg_bytes, const uint32_t G, const uint64_t n_parts, const uint32_t mis,
    const uint64_t len, const FindAllBufs fb, const Chain2Bufs cb, const int pass, int* err) {
  __shared__ uint2 skeys[CHAIN2_WARPS][32][CH2 + 1];
  __shared__ uint32_t sreps[CHAIN2_WARPS][32][CH2 + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t p0 = ((uint64_t)blockIdx.x * CHAIN2_WARPS + warp) * 32;
  if (p0 >= n_parts) return;
  const uint64_t p = p0 + lane;
  const bool live = p < n_parts;
  long long cursor = 0;
  if (live) {
    if (p == 0) cursor = 0;
    else if (pass == 0) cursor = (long long)(p * G * seg_bytes) - (long long)mis;
    else cursor = cb.exit_prev[p - 1];
    if (cursor < 0) cursor = 0;
  }
  unsigned long long nsel = 0, nreps = 0;
  for (uint32_t g = 0; g < G; g++) {
    // segment g of every lane's part
    const uint64_t seg = p * G + g;
    const bool has = live && seg < n_seg;
    const uint32_t c = has ? fb.count[seg] : 0;
    const long long seg_pos = (long long)(seg * seg_bytes) - (long long)mis;
    if (has) { cb.seg_sel[seg] = (uint32_t)nsel; cb.seg_reps[seg] = nreps; }
    uint32_t cmax = c;
#pragma unroll
    for (int o = 16; o; o >>= 1) cmax = max(cmax, __shfl_xor_sync(0xFFFFFFFFu, cmax, o));
    for (uint32_t r0 = 0; r0 < cmax; r0 += CH2) {
      // stage: for each lane-part j, the warp reads records r0..r0+31 of its segment (256 contiguous bytes)
      for (int j = 0; j < 32; j++) {
        const uint32_t cj = __shfl_sync(0xFFFFFFFFu, c, j);
        if (r0 < cj) {
          const uint64_t segj = (p0 + j) * G + g;
          if (r0 + lane < cj) skeys[warp][j][lane] = fb.keys[segj * fb.K + r0 + lane];
        }
      }
      __syncwarp();
      const uint32_t hi = c > r0 ? min(c - r0, (uint32_t)CH2) : 0;
      for (uint32_t r = 0; r < hi; r++) {
        const uint2 k = skeys[warp][lane][r];
        uint32_t reps = 0;
        if (k.y != KEY_INVALID) {
          const long long s = seg_pos + (long long)k.x;
          if (s >= cursor && (unsigned long long)cursor < len) {
            if (ENGINE == FIND_TDFA) {
              // offset += len(match) from the SLICE start (compiler.go:630-636): the record at s is
              // returned once per cursor value o, o+L, ... <= s
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
          }
        }
        sreps[warp][lane][r] = reps;
      }
      __syncwarp();
      // write the chunk's repeat counts back, coalesced
      for (int j = 0; j < 32; j++) {
        const uint32_t cj = __shfl_sync(0xFFFFFFFFu, c, j);
        if (r0 + lane < cj) {
          const uint64_t segj = (p0 + j) * G + g;
          fb.reps[segj * fb.K + r0 + lane] = sreps[warp][j][lane];
        }
      }
      __syncwarp();
    }
  }
  if (live) {
    cb.exit_cur[p] = cursor;
    if (pass > 0 && cb.exit_prev[p] != cursor) cb.changed[pass & 63] = 1;
    cb.part_sel[p] = nsel;
    cb.part_reps[p] = nreps;
  }
}

// ---- ordered output, one warp per segment -----------------------------------------------------------
// n_limit < 0: everything; otherwise the expanded list is cut after n_limit matches (a record's reps are
// clipped, later records dropped).  *n_written = number of records kept (written by whoever holds the last).
template <int ENGINE>
__global__ void findall_emit2_kernel(const DevMeta m, const uint64_t n_seg, const uint32_t seg_bytes, const uint32_t G,
                                     const uint32_t mis, const uint64_t len, const FindAllBufs fb, const Chain2Bufs cb,
                                     const unsigned long long* __restrict__ sel_base, const unsigned long long* __restrict__ reps_base,
                                     const unsigned long long* __restrict__ totals, const long long n_limit,
                                     int64_t* __restrict__ out, uint32_t* __restrict__ out_reps, const uint64_t cap_records,
                                     unsigned long long* n_written) {
  const uint64_t seg = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (seg == 0 && lane == 0 && (n_limit < 0 || totals[1] <= (unsigned long long)n_limit)) *n_written = totals[0];
  if (seg >= n_seg) return;
  const uint32_t c = fb.count[seg];
  if (c == 0) return;
  const uint64_t p = seg / G;
  unsigned long long o = sel_base[p] + cb.seg_sel[seg];
  unsigned long long cum = reps_base[p] + cb.seg_reps[seg];
  if (n_limit >= 0 && cum >= (unsigned long long)n_limit) return;
  const int nc = ENGINE == FIND_TDFA ? m.t_ntags : m.num_cap;
  const long long seg_pos = (long long)(seg * seg_bytes) - (long long)mis;
  for (uint32_t r0 = 0; r0 < c; r0 += 32) {
    const uint32_t r = r0 + lane;
    uint32_t reps = r < c ? fb.reps[seg * fb.K + r] : 0;
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, reps != 0);
    if (bal == 0) continue;
    const uint32_t before = __popc(bal & ((1u << lane) - 1u));
    unsigned long long incl = reps;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += y; }
    const unsigned long long my_cum = cum + incl - reps;
    if (reps != 0) {
      bool keep = true;
      const unsigned long long idx = o + before;
      if (n_limit >= 0) {
        if (my_cum >= (unsigned long long)n_limit) keep = false;
        else if (my_cum + reps >= (unsigned long long)n_limit) {
          reps = (uint32_t)((unsigned long long)n_limit - my_cum);
          *n_written = idx + 1;   // this record is the last one kept
        }
      }
      if (keep && idx < cap_records) {
        const uint64_t rr = seg * fb.K + r;
        const uint2 k = fb.keys[rr];
        const long long s0 = seg_pos + (long long)k.x, s = s0 + fb.out_base, e = s + (long long)k.y;
        int64_t* dst = out + idx * (uint64_t)nc;
        dst[0] = s; dst[1] = e;
        for (int g = 1; g < nc / 2; g++) {
          const int32_t a = fb.caps[rr * fb.cw + 2 * g - 2], b = fb.caps[rr * fb.cw + 2 * g - 1];
          if (ENGINE == FIND_TDFA) {
            if (a >= 0) { dst[2 * g] = s + a; dst[2 * g + 1] = s + b; } else { dst[2 * g] = -1; dst[2 * g + 1] = -1; }
          } else {
            const long long av = a == CAP_ZERO ? 0 : s + a, bv = b == CAP_ZERO ? 0 : s + b;
            if (av <= bv && bv <= (long long)len) { dst[2 * g] = av; dst[2 * g + 1] = bv; } else { dst[2 * g] = -1; dst[2 * g + 1] = -1; }
          }
        }
        out_reps[idx] = reps;
      }
    }
    o += __popc(bal);
    cum += __shfl_sync(0xFFFFFFFFu, incl, 31);
  }
}

}  // namespace rgx
