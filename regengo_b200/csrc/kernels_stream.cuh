// stream.FindReader (internal/compiler/streaming.go:85-255) for a reader that fills every Read.
//
// With full reads the reference's loop is a data-independent chunk schedule: chunk k
// is stream[k*(B-L), k*(B-L)+B) and it reports the matches of repeated FindBytesReuse calls on
// chunk[searchPos:] whose end is <= B-L (the tail chunk reports everything; if the last read filled
// the buffer exactly, one more "flush" pass runs over the final L bytes).  Chunks are independent,
// so one device thread replays one chunk exactly as the generated loop does: the skip-restart rule
// inside FindBytesReuse (Q1), re-anchoring at every searchPos (Q3), and locating the match text
// with bytes.Index from searchPos (Q16) all fall out of running the same steps in the same order.
// Two passes: count per chunk, exclusive scan, then the same replay writes at its offset.
#pragma once
#include "engines.cuh"
#include "kernels_batch.cuh"

namespace rgx {

struct ChunkPlan {
  uint64_t stride;        // B - L
  uint64_t B, L;
  uint64_t total_len;     // whole stream
  uint64_t n_chunks;      // including the flush pass, if any
  uint64_t flush_chunk;   // index of the flush pass (== n_chunks when there is none)
  uint64_t flush_start;   // stream offset of the flush pass
};

// geometry of chunk k: stream offset, length, and whether the deferral rule applies
__device__ __forceinline__ void chunk_geometry(const ChunkPlan& cp, uint64_t k, uint64_t& start, uint64_t& dlen, bool& full) {
  if (k == cp.flush_chunk) { start = cp.flush_start; dlen = cp.total_len - cp.flush_start; full = false; return; }
  start = k * cp.stride;
  const uint64_t avail = cp.total_len - start;
  dlen = avail < cp.B ? avail : cp.B;
  // isFull := n == BufferSize - leftover: the read filled the buffer (streaming.go:177)
  full = avail >= cp.B;
}

// bytes.Index(hay[from:], needle) with needle = hay[ns:ns+nl], known to occur at ns >= from
__device__ __forceinline__ int64_t index_of_text(const uint8_t* hay, int64_t from, int64_t ns, int64_t nl) {
  if (nl == 0) return from;
  for (int64_t q = from; q < ns; q++) {
    int64_t j = 0;
    while (j < nl && hay[q + j] == hay[ns + j]) j++;
    if (j == nl) return q;
  }
  return ns;
}

// MODE 0: count matches per chunk.  MODE 1: write them at base[chunk].
template <int MODE>
__global__ void __launch_bounds__(128) find_reader_kernel(const DevMeta m, const uint32_t* __restrict__ gimg, const int in_smem,
                                                          const uint8_t* __restrict__ d_stream, const uint64_t base_off,
                                                          const ChunkPlan cp, const uint64_t first_chunk, const uint64_t n_run,
                                                          unsigned long long* __restrict__ counts,
                                                          const unsigned long long* __restrict__ bases, int64_t* __restrict__ out_soff,
                                                          int32_t* __restrict__ out_chunk, int64_t* __restrict__ out_rec,
                                                          const uint64_t cap, const ScratchPlan sp, int* err) {
  extern __shared__ __align__(16) uint32_t smem_img[];
  __shared__ __align__(8) unsigned long long mbar;
  const uint32_t* img = gimg;
  if (in_smem) { stage_image_tma(smem_img, gimg, m.image_words, &mbar); img = smem_img; }
  Scratch sc;
  sc.stack = sp.stack; sc.cstack = sp.cstack; sc.visited = sp.visited;
  sc.stack_cap = sp.stack_cap; sc.cstack_cap = sp.cstack_cap; sc.visited_words = sp.visited_words;
  sc.stride = sp.stride; sc.tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nc = m.find_engine == FIND_TDFA ? m.t_ntags : m.num_cap;

  for (uint64_t j = sc.tid; j < n_run; j += sp.stride) {
    const uint64_t k = first_chunk + j;
    uint64_t cstart, dlen;
    bool full;
    chunk_geometry(cp, k, cstart, dlen, full);
    const uint8_t* chunk = d_stream + (cstart - base_off);
    const int64_t data_len = (int64_t)dlen;
    int64_t search_pos = 0;
    unsigned long long n = 0;
    unsigned long long w = MODE == 1 ? bases[j] : 0;
    int64_t rec[MAX_CAPS];
    while (search_pos < data_len) {
      int found;
      if (m.find_engine == FIND_TDFA) {
        found = tdfa_find(m, img, chunk + search_pos, data_len - search_pos, 0, rec);
      } else {
        int32_t caps[MAX_CAPS];
        found = bt_machine<MODE_FIND>(m, img, chunk + search_pos, data_len - search_pos, 0, caps, sc, err);
        if (found) bt_emit_record(caps, nc, 0, data_len - search_pos, 0, rec);
      }
      if (!found) break;
      const int64_t mlen = rec[1] - rec[0];
      const int64_t mstart = index_of_text(chunk, search_pos, search_pos + rec[0], mlen);
      const int64_t mend = mstart + mlen;
      if (full && mend > data_len - (int64_t)cp.L) break;  // too close to the boundary: next chunk's job
      if (MODE == 1 && w < cap) {
        out_soff[w] = (int64_t)cstart + mstart;
        out_chunk[w] = (int32_t)k;
        for (int g = 0; g < nc; g++) out_rec[w * nc + g] = rec[g] < 0 ? -1 : (int64_t)cstart + search_pos + rec[g];
      }
      w++; n++;
      if (mlen > 0) search_pos = mend; else search_pos++;
    }
    if (MODE == 0) counts[j] = n;
  }
}

__global__ void exclusive_scan_u64_kernel(const uint64_t n, const unsigned long long* __restrict__ in, unsigned long long* out,
                                          unsigned long long* total) {
  __shared__ unsigned long long s[1024];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint64_t base = 0; base < n; base += 1024) {
    const uint64_t i = base + threadIdx.x;
    const unsigned long long a = i < n ? in[i] : 0;
    s[threadIdx.x] = a;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      unsigned long long x = 0;
      if ((int)threadIdx.x >= o) x = s[threadIdx.x - o];
      __syncthreads();
      s[threadIdx.x] += x;
      __syncthreads();
    }
    if (i < n) out[i] = carry + s[threadIdx.x] - a;
    __syncthreads();
    if (threadIdx.x == 1023) carry += s[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

// ---------------------------------------------------------------------------------------------------
// Fast FindReader for backtracking patterns (no EmptyWidth, not anchored, not memoised, not nullable).
//
// FindBytesReuse on chunk[searchPos:] is a chain of attempts: the attempt at a either matches or fails at
// an offset f and the next attempt starts at f + 1.  Whether the attempt at a matches, its
// length and f do not depend on searchPos, and they do not depend on where the chunk ends as long as the
// attempt never looked at or past that end.  So:
//   TABLE   one attempt per stream position, all positions in parallel: a 16-bit entry
//           {match:1 | reach:7 | len-or-next:8} -- reach = how far the attempt looked, len = match length or
//           next = distance to the next attempt (for a byte that cannot start a match: to the next byte
//           that can).  Values that do not fit mean "decide inline".
//   CHASE   one warp per chunk replays the reference loop over the table: follow next-pointers to the first
//           match (inline exact attempt with the chunk's limit when an entry is marked or reaches the chunk
//           end), locate the match text with bytes.Index from searchPos (Q16), apply the deferral rule,
//           move searchPos; lists {searchPos, attempt start, text position}.  Tiles of the chunk are chased by
//           all lanes in parallel from sync points (see below).
//   RECORDS one lane per listed match re-runs that single attempt on chunk[searchPos:] for the captures.
// Every step is the reference's own; only the order of evaluation differs.
constexpr uint32_t RT_SLOW_REACH = 127, RT_SLOW_VAL = 255;
constexpr int RT_TILE = 2048;   // positions per tile of the chase (one warp; 64 per lane)

__device__ __forceinline__ Scratch scratch_of(const ScratchPlan& sp) {
  Scratch sc;
  sc.stack = sp.stack; sc.cstack = sp.cstack; sc.visited = sp.visited;
  sc.stack_cap = sp.stack_cap; sc.cstack_cap = sp.cstack_cap; sc.visited_words = sp.visited_words;
  sc.stride = sp.stride; sc.tid = blockIdx.x * blockDim.x + threadIdx.x;
  return sc;
}

// One TDFA start without tags (tdfa.go:937-975 minus the tag bookkeeping): last accepting offset or -1, and the
// reach (one past the byte that stopped the walk; l + 1 when the walk ran into the end of the input, because
// then the outcome may depend on l).  For tables without end-of-text accepts and with one start state.
__device__ __forceinline__ int64_t tdfa_walk_light(const DevMeta& m, const uint32_t* __restrict__ img, const uint8_t* __restrict__ in,
                                                   const int64_t l, const int64_t start, int64_t* reach) {
  const uint32_t* trans = img + m.off_t_trans;
  const uint32_t* acc = img + m.off_t_accept;
  uint32_t state = (uint32_t)m.t_start_any;
  int64_t match_end = -1;
  for (int64_t i = start; i < l; i++) {
    const uint32_t c = in[i];
    if (c >= 128) { *reach = i + 1; return match_end; }
    const uint32_t nx = trans[state * 128 + c] & 0xFFFFu;
    if (nx == TDFA_NONE) { *reach = i + 1; return match_end; }
    state = nx;
    if (acc[state] & 1u) match_end = i + 1;
  }
  *reach = l + 1;
  return match_end;
}

// One attempt at position s of in[0:l): 1 = match (*mlen), 0 = fail (*next_start = where FindBytes tries next:
// failure offset + 1 for the goto-machine, s + 1 for the TDFA, which tries every start).
__device__ __forceinline__ int reader_attempt(const DevMeta& m, const uint32_t* __restrict__ img, const uint8_t* __restrict__ in,
                                              const int64_t l, const int64_t s, const Scratch& sc, int* err, int64_t* mlen,
                                              int64_t* next_start, int64_t* reach) {
  if (m.find_engine == FIND_TDFA) {
    const int64_t me = tdfa_walk_light(m, img, in, l, s, reach);
    if (me >= 0) { *mlen = me - s; return 1; }
    *next_start = s + 1;
    return 0;
  }
  int32_t caps[MAX_CAPS];
  int64_t f = 0;
  if (bt_machine<MODE_FINDALL, true>(m, img, in, l, s, caps, sc, err, -1, 0, 0, &f, reach)) { *mlen = caps[1]; return 1; }
  *next_start = f + 1;
  return 0;
}

// d_stream[0:len) = the shard; entry i describes the attempt at shard position i with the whole shard visible
__global__ void __launch_bounds__(256) find_reader_table_kernel(const DevMeta m, const uint32_t* __restrict__ gimg, const int in_smem,
                                                                const uint8_t* __restrict__ d_stream, const uint64_t len,
                                                                uint16_t* __restrict__ table, const ScratchPlan sp, int* err,
                                                                int* slow_flag) {
  extern __shared__ __align__(16) uint32_t smem_img[];
  __shared__ __align__(8) unsigned long long mbar;
  const uint32_t* img = gimg;
  if (in_smem) { stage_image_tma(smem_img, gimg, m.image_words, &mbar); img = smem_img; }
  const Scratch sc = scratch_of(sp);
  const uint32_t* first = img + m.off_first;
  bool any_slow = false;
  const uint64_t n_units = (len + 63) / 64;     // 64 consecutive positions per thread
  for (uint64_t u = sc.tid; u < n_units; u += sp.stride) {
    const uint64_t p0 = u * 64;
    const uint32_t nb = (uint32_t)min((uint64_t)64, len - p0);
    // which of my bytes can start a match (16-byte loads when the unit is whole and aligned)
    unsigned long long cand = 0;
    const bool whole = nb == 64 && (((uintptr_t)(d_stream + p0)) & 15u) == 0;
    if (whole) {
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const uint4 v = *reinterpret_cast<const uint4*>(d_stream + p0 + 16 * q);
        const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int b = 0; b < 16; b++) {
          const uint32_t c = (ws[b >> 2] >> (8 * (b & 3))) & 255u;
          cand |= (unsigned long long)((first[c >> 5] >> (c & 31)) & 1u) << (16 * q + b);
        }
      }
    } else {
      for (uint32_t j = 0; j < nb; j++) {
        const uint32_t c = d_stream[p0 + j];
        if ((first[c >> 5] >> (c & 31)) & 1u) cand |= 1ull << j;
      }
    }
    // the attempts (one per candidate), kept in a small per-thread array
    uint16_t ent[64];
    for (unsigned long long rest = cand; rest;) {
      const uint32_t j = (uint32_t)__ffsll((long long)rest) - 1;
      rest &= rest - 1;
      int64_t ml = 0, ns = 0, reach = 0;
      const int64_t s = (int64_t)(p0 + j);
      const int ok = reader_attempt(m, img, d_stream, (int64_t)len, s, sc, err, &ml, &ns, &reach);
      const int64_t rr = reach - s;
      const uint32_t r7 = rr >= (int64_t)RT_SLOW_REACH || rr < 1 ? RT_SLOW_REACH : (uint32_t)rr;
      uint32_t entry;
      if (ok) {
        entry = 0x8000u | (r7 << 8) | (ml >= (int64_t)RT_SLOW_VAL || ml < 0 ? RT_SLOW_VAL : (uint32_t)ml);
      } else {
        const int64_t nx = ns - s;
        entry = (r7 << 8) | (nx >= (int64_t)RT_SLOW_VAL || nx < 1 ? RT_SLOW_VAL : (uint32_t)nx);
      }
      ent[j] = (uint16_t)entry;
      any_slow = any_slow || ((entry >> 8) & 0x7Fu) == RT_SLOW_REACH || (entry & 0xFFu) == RT_SLOW_VAL;
    }
    auto entry_of = [&](const uint32_t j) -> uint32_t {
      if ((cand >> j) & 1ull) return ent[j];
      const unsigned long long rest = j + 1 < 64 ? cand >> (j + 1) : 0ull;
      uint32_t dist = rest ? (uint32_t)__ffsll((long long)rest) : 64u - j;   // next candidate, or my last byte + 1
      if (j + dist > nb) dist = nb - j;
      return (1u << 8) | dist;                                                  // reach 1, next = dist (<= 64)
    };
    if (whole && (((uintptr_t)(table + p0)) & 15u) == 0) {
#pragma unroll
      for (int q = 0; q < 8; q++) {
        uint4 v;
        v.x = entry_of(8 * q) | (entry_of(8 * q + 1) << 16);
        v.y = entry_of(8 * q + 2) | (entry_of(8 * q + 3) << 16);
        v.z = entry_of(8 * q + 4) | (entry_of(8 * q + 5) << 16);
        v.w = entry_of(8 * q + 6) | (entry_of(8 * q + 7) << 16);
        *reinterpret_cast<uint4*>(table + p0 + 8 * q) = v;
      }
    } else {
      for (uint32_t j = 0; j < nb; j++) table[p0 + j] = (uint16_t)entry_of(j);
    }
  }
  if (any_slow) *slow_flag = 1;   // some entry says "decide inline": the chase keeps to the sequential form
}

// Bit planes of one unit of 64 attempt positions of a STRAIGHT-LINE program from the unit's 96 bytes (bw[24], bytes
// that do not exist = 0xFF): cand = positions whose byte passes step 0, alive = positions that pass all S steps (the
// matches, length S), pl[0..5] = the number of steps survived, bit-sliced (an attempt that survives cnt < S steps
// fails at offset cnt: the next attempt starts cnt + 1 further, and it looked at cnt + 1 bytes).
__device__ __forceinline__ void linear_unit_planes(const DevMeta& m, const int S, const uint8_t* cm, uint32_t* bw, unsigned long long& cand,
                                                   unsigned long long& alive, unsigned long long* pl) {
  const int ncls = m.sl_ncls;   // (S = m.sl_n for a whole straight-line program, m.slp_n for a straight-line prefix)
#pragma unroll
  for (int q = 0; q < 24; q++) {
    const uint32_t x = bw[q];
    bw[q] = (uint32_t)cm[x & 255u] | ((uint32_t)cm[(x >> 8) & 255u] << 8) | ((uint32_t)cm[(x >> 16) & 255u] << 16) |
            ((uint32_t)cm[x >> 24] << 24);
  }
  // per class: a 96-bit mask (lo: positions 0..63, hi: 64..95)
  unsigned long long mlo[8];
  uint32_t mhi[8];
  for (int k = 0; k < ncls; k++) {
    unsigned long long lo = 0;
    uint32_t hi = 0;
#pragma unroll
    for (int q = 0; q < 24; q++) {
      const uint32_t t = (bw[q] >> k) & 0x01010101u;                    // bit 0 of each byte
      const uint32_t nib = ((t * 0x10204080u) >> 28) & 0xFu;            // -> 4 consecutive bits, byte 0 lowest
      if (q < 16) lo |= (unsigned long long)nib << (4 * q); else hi |= nib << (4 * (q - 16));
    }
    mlo[k] = lo; mhi[k] = hi;
  }
  alive = ~0ull; cand = 0;
  unsigned long long pl0 = 0, pl1 = 0, pl2 = 0, pl3 = 0, pl4 = 0, pl5 = 0;
  for (int i = 0; i < S; i++) {
    const int k = m.sl_cls[i];
    const unsigned long long sh = i == 0 ? mlo[k] : ((mlo[k] >> i) | ((unsigned long long)mhi[k] << (64 - i)));
    alive &= sh;
    if (i == 0) cand = alive;
    unsigned long long carry = alive, t;
    t = pl0 & carry; pl0 ^= carry; carry = t;
    t = pl1 & carry; pl1 ^= carry; carry = t;
    t = pl2 & carry; pl2 ^= carry; carry = t;
    t = pl3 & carry; pl3 ^= carry; carry = t;
    t = pl4 & carry; pl4 ^= carry; carry = t;
    pl5 ^= carry;
  }
  pl[0] = pl0; pl[1] = pl1; pl[2] = pl2; pl[3] = pl3; pl[4] = pl4; pl[5] = pl5;
}

struct ReaderHit { long long search_abs; uint32_t d_true, d_text; unsigned long long chunk; };   // stream offset of searchPos; attempt start / text position relative to it; ChunkIndex

// bytes.Index(hay[from:], hay[ns:ns+nl]) by a whole warp: lane i tests from + i, from + 32 + i, ...
__device__ __forceinline__ int64_t warp_index_of_text(const uint8_t* hay, int64_t from, int64_t ns, int64_t nl, int lane) {
  if (nl == 0) return from;
  for (int64_t b = from; b < ns; b += 32) {
    const int64_t q = b + lane;
    bool eq = q < ns;
    for (int64_t j = 0; eq && j < nl; j++) eq = hay[q + j] == hay[ns + j];
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, eq);
    if (bal) return b + (__ffs(bal) - 1);
  }
  return ns;
}

// bytes.Index(hay[from:], hay[ns:ns+nl]) by ONE lane, testing only positions the table does not rule out: an entry
// {no match, reach 1, next > 1} is a byte outside the first-byte set, and so are the next - 1 bytes after it; the match
// text starts with a byte of that set.
__device__ __forceinline__ int64_t table_index_of_text(const uint8_t* hay, const uint16_t* tab, int64_t from, int64_t ns, int64_t nl) {
  for (int64_t q = from; q < ns;) {
    const uint32_t e = tab[q];
    if (!(e & 0x8000u) && ((e >> 8) & 0x7Fu) == 1u && (e & 0xFFu) > 1u) { q += (int64_t)(e & 0xFFu); continue; }
    int64_t j = 0;
    while (j < nl && hay[q + j] == hay[ns + j]) j++;
    if (j == nl) return q;
    q++;
  }
  return ns;
}

// SYNC POINTS (what lets the lanes of a warp chase one chunk in parallel).  Call extent(a) = max(reach, len-or-next) of
// the attempt at a: it looks no further than a + extent(a), and whatever the replay does at a -- fail and restart, match
// and continue behind the match, match with the text located earlier (Q16) -- its next position is at most
// a + extent(a).  If every a < c has a + extent(a) <= c, then c is visited by EVERY replay that starts before c: take
// the last visited position v < c; the next one is >= c by choice of v and <= v + extent(v) <= c.  From c on the replay
// depends on the past only through searchPos (the end of the last match), which matters for one thing: where
// bytes.Index starts looking for the text of the next match.  The tile chase in find_reader_chase_kernel is built on
// this; an over-estimated cover (max of a + extent(a)) only makes it find fewer sync points, never a wrong one.

// does hay[q:q+nl] equal hay[np:np+nl]?
__device__ __forceinline__ bool same_text(const uint8_t* hay, int64_t q, int64_t np, int64_t nl) {
  int64_t j = 0;
  while (j < nl && hay[q + j] == hay[np + j]) j++;
  return j == nl;
}

// MODE 0: count per chunk.  MODE 1: list the hits at bases[chunk].  MODE 2: one pass -- list the hits of chunk j
// at j * region (a chunk cannot hold more than `region` matches) and count them; a chunk that would overflow
// its region sets ERR_SLAB and the caller falls back to the two-pass form.
// One WARP per chunk: tiles of RT_TILE positions chased by all lanes at once (below); the last 127 positions of a chunk,
// short chunks and tables with "decide inline" entries take the sequential form, where the lanes fetch 32 table entries
// at a time and the attempt chain is followed through them with shuffles.  LINEAR: a straight-line program
// (device_program.cu: sl_*) -- no table at all, the attempts of a tile are bit planes computed from its bytes.
template <int MODE, bool LINEAR>
__global__ void __launch_bounds__(128) find_reader_chase_kernel(const DevMeta m, const uint32_t* __restrict__ gimg, const int in_smem,
                                                                const uint8_t* __restrict__ d_stream, const uint64_t base_off,
                                                                const uint16_t* __restrict__ table, const ChunkPlan cp,
                                                                const uint64_t first_chunk, const uint64_t n_run,
                                                                unsigned long long* __restrict__ counts,
                                                                const unsigned long long* __restrict__ bases, ReaderHit* __restrict__ hits,
                                                                const uint64_t cap, const ScratchPlan sp, int* err, const int* slow_flag) {
  extern __shared__ __align__(16) uint32_t smem_img[];
  __shared__ __align__(8) unsigned long long mbar;
  const uint32_t* img = gimg;
  if (in_smem) { stage_image_tma(smem_img, gimg, m.image_words, &mbar); img = smem_img; }
  const Scratch sc = scratch_of(sp);
  const int lane = threadIdx.x & 31;
  __shared__ __align__(16) uint16_t Etile[4][RT_TILE];
  __shared__ unsigned long long Mtile[4][32];
  const bool no_slow = LINEAR || *slow_flag == 0;
  for (uint64_t j = sc.tid >> 5; j < n_run; j += sp.stride >> 5) {
    const uint64_t k = first_chunk + j;
    uint64_t cstart, dlen;
    bool full;
    chunk_geometry(cp, k, cstart, dlen, full);
    const uint8_t* chunk = d_stream + (cstart - base_off);
    const uint16_t* tab = table + (cstart - base_off);
    const int64_t data_len = (int64_t)dlen;
    int64_t pos = 0;
    unsigned long long n = 0, w = MODE == 1 ? bases[j] : MODE == 2 ? j * cap : 0;
    const unsigned long long wend = MODE == 2 ? w + cap : cap;
    // ---- TILE CHASE: the chunk in tiles of RT_TILE positions, entries staged in shared memory ----
    // Warp-uniform state: pos = searchPos, cur = position of the next attempt, cover = max over every position a' before
    // the tile of a' + extent(a') (exact, or an over-estimate: that only delays sync points).  Per tile:
    //   1. one coalesced load of the tile's entries; lane i owns the unit of 64 positions [P + 64 i, P + 64 i + 64);
    //   2. unit reach m_i = max (a + extent(a)) over its entries, exclusive prefix maximum over the lanes -> the cover
    //      entering each unit -> the unit's FIRST SYNC POINT (see above: a position every replay from before it visits);
    //      the lane that owns cur starts there instead;
    //   3. every lane with a start chases the entries up to the next lane's start (it must arrive exactly there) -- all
    //      lanes at once, a handful of shared-memory reads each; counts and last match ends give every match its
    //      searchPos (the end of the match before it, in order);
    //   4. the same chase again, now testing each match the way the reference does: text located by bytes.Index from
    //      searchPos (Q16) and the deferral rule.  A match that is deferred ends the chunk; a match whose text occurs
    //      earlier is an EVENT: everything before the first event in order is final and written out, then the event is
    //      reported where bytes.Index finds it, searchPos moves behind that copy and the tiles restart there (with the
    //      conservative cover cur + 254), exactly as the reference searches on.
    // Attempts at positions >= data_len - 127 may look past the chunk end: the sequential form below takes over there.
    int64_t cur = 0;
    bool stop = false;
    if (LINEAR && data_len >= 2 * RT_TILE && data_len < (1ll << 30)) {
      // ---- the same for a STRAIGHT-LINE program, without a table: every lane computes the bit planes of its unit from
      // the unit's 96 bytes (linear_unit_planes) and the chase reads the attempt at a position off them: match iff the
      // `alive` bit is set (length S), else next attempt cnt + 1 further (cnt = steps survived; 0 for a byte that cannot
      // start a match).  Cover uses the bound S for every candidate (an over-estimate only delays sync points).
      unsigned long long* U = reinterpret_cast<unsigned long long*>(Etile[threadIdx.x >> 5]);   // U[plane * 32 + unit]: cand, alive, pl0..pl5
      const uint8_t* cm = reinterpret_cast<const uint8_t*>(img + m.off_sl_cm);
      const int S = m.sl_n;
      const int dl = (int)data_len;
      const int lim = dl - 127;
      const int P0 = -(int)((uintptr_t)chunk & 15u);        // unit bases keep the 16-byte alignment of the byte loads
      int P = P0, cover = 0, cu = 0, sp_w = 0;
      const int defer_line = dl - (int)cp.L;
      auto steps_at = [&](const int u, const int b) -> int {   // steps survived by the attempt at unit u, bit b
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) cnt |= (int)((U[(2 + i) * 32 + u] >> b) & 1ull) << i;
        return cnt;
      };
      // first copy of the text at ta in [from, ta), or ta: a copy of a match's text is itself a match entry
      unsigned long long* Mprev = Mtile[threadIdx.x >> 5];   // `alive` of the tile before this one (when prev_ok)
      bool prev_ok = false;
      auto index_of = [&](const int from, const int ta) -> int {
        if (from >= ta) return ta;
        if (from < P && !(prev_ok && from >= P - RT_TILE)) return (int)index_of_text(chunk, from, ta, S);
        // units are numbered from the previous tile's first: 0..31 previous, 32..63 this tile
        const int base = P - RT_TILE;
        for (int u = (from - base) >> 6; u <= (ta - 1 - base) >> 6; u++) {
          unsigned long long mk = u < 32 ? Mprev[u] : U[u];   // (U[32 + unit] = alive)
          const int ub0 = base + 64 * u;
          if (from > ub0) mk &= ~0ull << (from - ub0);
          if (ta < ub0 + 64) mk &= (1ull << (ta - ub0)) - 1ull;
          while (mk) {
            const int q = ub0 + __ffsll((long long)mk) - 1;
            mk &= mk - 1;
            if (same_text(chunk, q, ta, S)) return q;
          }
        }
        return ta;
      };
      // one step of the replay at position a (< limit): returns the next position, *hit = a match starts at a
      auto step_at = [&](const int a, const int limit, bool* hit) -> int {
        const int u = (a - P) >> 6, b = (a - P) & 63;
        const unsigned long long cu_ = U[u] >> b;
        *hit = false;
        if (!(cu_ & 1ull)) {
          // bytes that cannot start a match: on to the next one that can (the reference tries each and fails at once)
          const int nx = cu_ ? a + __ffsll((long long)cu_) - 1 : P + 64 * (u + 1);
          return min(nx, limit);
        }
        if ((U[32 + u] >> b) & 1ull) { *hit = true; return a + S; }
        return a + steps_at(u, b) + 1;
      };
      while (!stop && cu < lim) {
        const int tile_end = min(P + RT_TILE, dl);
        const int ub = P + 64 * lane, ue = min(ub + 64, tile_end);
        // 1. planes of my unit
        {
          uint32_t bw[24];
          if (ub >= 0 && ub + 96 <= dl) {
#pragma unroll
            for (int q = 0; q < 6; q++) {
              const uint4 v = *reinterpret_cast<const uint4*>(chunk + ub + 16 * q);
              bw[4 * q] = v.x; bw[4 * q + 1] = v.y; bw[4 * q + 2] = v.z; bw[4 * q + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int q = 0; q < 24; q++) {
              uint32_t x = 0;
              for (int b = 0; b < 4; b++) {
                const int pp = ub + 4 * q + b;
                x |= (uint32_t)((pp >= 0 && pp < dl) ? chunk[pp] : 0xFFu) << (8 * b);
              }
              bw[q] = x;
            }
          }
          unsigned long long cand, alive, pl[6];
          linear_unit_planes(m, S, cm, bw, cand, alive, pl);
          if (ub >= tile_end) { cand = 0; alive = 0; }
          __syncwarp();
          U[lane] = cand; U[32 + lane] = alive;
#pragma unroll
          for (int i = 0; i < 6; i++) U[(2 + i) * 32 + lane] = pl[i];
        }
        __syncwarp();
        // 2. cover entering the unit, start of the lane's chase
        const unsigned long long myc = U[lane];
        const int mreach = myc ? ub + 63 - __clzll((long long)myc) + S : 0;
        int incl = mreach;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl = max(incl, y); }
        const int before = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
        const int cin = lane == 0 ? cover : max(cover, before);
        const int lane_cur = (cu - P) >> 6;
        int start = -1;
        if (lane == lane_cur) start = cu;
        else if (lane > lane_cur && ub < min(tile_end, lim)) {
          // positions of my unit with no candidate less than S before them (inside the unit) and not below the cover
          unsigned long long blocked = myc << 1;
          for (int sh = 1; sh < S - 1; sh <<= 1) blocked |= blocked << min(sh, S - 1 - sh);
          unsigned long long ok = ~blocked;
          if (cin > ub) ok = cin - ub >= 64 ? 0ull : ok & (~0ull << (cin - ub));
          const int top = min(ue, lim) - ub;
          if (top < 64) ok &= (1ull << top) - 1ull;
          if (ok) start = ub + __ffsll((long long)ok) - 1;
        }
        const uint32_t startm = __ballot_sync(0xFFFFFFFFu, start >= 0);
        const uint32_t higher = lane == 31 ? 0u : (startm >> (lane + 1)) << (lane + 1);
        const int nstart = __shfl_sync(0xFFFFFFFFu, start, higher ? __ffs(higher) - 1 : lane);
        const int e_end = higher ? nstart : min(tile_end, lim);
        // 3. ONE chase per lane: matches, last match end, the first four match positions (kept in registers), and the
        //    reference's tests for every match but the lane's first -- their searchPos is the end of the match before
        //    them, in the same lane.  (ev: 1 = deferred, the chunk ends; 2 = the text occurs earlier)
        uint32_t n_l = 0, k_l = 0;
        int last_mend = -1, a = start;
        int m0 = 0, m1 = 0, m2 = 0, m3 = 0;
        int ev = 0, ev_q = 0, ev_a = 0, ev_sp = 0;
        if (start >= 0)
          while (a < e_end && !ev) {
            bool hit;
            const int nx = step_at(a, e_end, &hit);
            if (hit) {
              if (n_l == 0) m0 = a; else if (n_l == 1) m1 = a; else if (n_l == 2) m2 = a; else if (n_l == 3) m3 = a;
              if (n_l) {
                const int q = index_of(last_mend, a);
                if (full && q + S > defer_line) ev = 1;
                else if (q < a) { ev = 2; ev_q = q; ev_a = a; ev_sp = last_mend; }
                else k_l++;
              }
              n_l++; last_mend = nx;
            }
            a = nx;
          }
        if (start >= 0 && higher && !ev && a != e_end) atomicOr(err, ERR_INTERNAL);   // (a sync point is visited by every replay)
        const int exit_a = a;
        int lm = last_mend;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xFFFFFFFFu, lm, o); if (lane >= o) lm = max(lm, y); }
        const int lm_before = __shfl_up_sync(0xFFFFFFFFu, lm, 1);
        const int sp_in = lane == 0 ? sp_w : max(sp_w, lm_before);
        const int lm_all = __shfl_sync(0xFFFFFFFFu, lm, 31);
        // 4. the lane's first match, now that its searchPos is known (the end of the last match of the lanes before)
        if (n_l) {
          const int q = index_of(sp_in, m0);
          if (full && q + S > defer_line) { ev = 1; k_l = 0; }
          else if (q < m0) { ev = 2; ev_q = q; ev_a = m0; ev_sp = sp_in; k_l = 0; }
          else k_l++;
        }
        const uint32_t evm = __ballot_sync(0xFFFFFFFFu, ev != 0);
        const int F = evm ? __ffs(evm) - 1 : 32;
        const uint32_t valid = lane < F ? n_l : lane == F ? k_l : 0u;
        uint32_t vincl = valid;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, vincl, o); if (lane >= o) vincl += y; }
        const uint32_t vtotal = __shfl_sync(0xFFFFFFFFu, vincl, 31);
        const int evF = __shfl_sync(0xFFFFFFFFu, ev, F & 31);
        const bool reloc = F < 32 && evF == 2;
        if (MODE == 2 && w + vtotal + (reloc ? 1u : 0u) > wend) { if (lane == 0) atomicOr(err, ERR_SLAB); n += vtotal + 1; stop = true; break; }
        if (MODE != 0 && valid) {
          unsigned long long wi = w + vincl - valid;
          if (valid <= 4) {
            // straight from the registers
            int sp = sp_in;
            for (uint32_t i = 0; i < valid; i++) {
              const int ma = i == 0 ? m0 : i == 1 ? m1 : i == 2 ? m2 : m3;
              if (wi < wend) {
                ReaderHit h;
                h.search_abs = (long long)cstart + sp; h.d_true = (uint32_t)(ma - sp); h.d_text = (uint32_t)(ma - sp); h.chunk = k;
                hits[wi] = h;
              }
              wi++;
              sp = ma + S;
            }
          } else {
            int sp = sp_in;
            uint32_t left = valid;
            a = start;
            while (left) {
              bool hit;
              const int nx = step_at(a, e_end, &hit);
              if (hit) {
                if (wi < wend) {
                  ReaderHit h;
                  h.search_abs = (long long)cstart + sp; h.d_true = (uint32_t)(a - sp); h.d_text = (uint32_t)(a - sp); h.chunk = k;
                  hits[wi] = h;
                }
                wi++; left--;
                sp = nx;
              }
              a = nx;
            }
          }
        }
        w += vtotal; n += vtotal;
        if (F == 32) {
          if (lm_all >= 0) sp_w = max(sp_w, lm_all);
          const int last = 31 - __clz((int)startm);
          cu = __shfl_sync(0xFFFFFFFFu, exit_a, last);
          cover = max(cover, __shfl_sync(0xFFFFFFFFu, incl, 31));
          P += RT_TILE;
          __syncwarp();                    // (the tests above read Mprev)
          Mprev[lane] = U[32 + lane];
          prev_ok = true;
        } else if (!reloc) {
          stop = true;     // too close to the boundary: the next chunk's job (and everything behind it)
        } else {
          const int q = __shfl_sync(0xFFFFFFFFu, ev_q, F), ta = __shfl_sync(0xFFFFFFFFu, ev_a, F), tsp = __shfl_sync(0xFFFFFFFFu, ev_sp, F);
          if (MODE != 0 && w < wend && lane == 0) {
            ReaderHit h;
            h.search_abs = (long long)cstart + tsp; h.d_true = (uint32_t)(ta - tsp); h.d_text = (uint32_t)(q - tsp); h.chunk = k;
            hits[w] = h;
          }
          w++; n++;
          sp_w = q + S; cu = sp_w;
          P = P0 + ((cu - P0) & ~63);
          cover = cu + 254;
          prev_ok = false;
        }
        __syncwarp();
      }
      pos = sp_w; cur = cu;
      if (!stop && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(const_cast<int*>(slow_flag)) + 2, 1ull);   // statistics: chunks on the tile chase
    } else if (!LINEAR && no_slow && data_len >= 2 * RT_TILE && data_len < (1ll << 30)) {
      // (positions are chunk-relative 32-bit integers in here)
      uint16_t* E = Etile[threadIdx.x >> 5];
      unsigned long long* MM = Mtile[threadIdx.x >> 5];   // match entries of each unit, one bit per position
      const int dl = (int)data_len;
      const int lim = dl - 127;
      const int P0 = -(int)(((uintptr_t)tab >> 1) & 7u);     // tile bases keep the 16-byte alignment of the table loads
      int P = P0, cover = 0, cu = 0, sp_w = 0;               // cu = cur, sp_w = pos (searchPos)
      const int defer_line = dl - (int)cp.L;
      // A straight-line program matches at q exactly when the bytes at q pass its S steps, so a copy of a match's text
      // IS a match entry: bytes.Index only has to look at the match entries between searchPos and the match.
      const bool copies_are_matches = m.sl_n > 0;
      auto hop_of = [](const uint32_t e) -> int {   // a run of bytes that cannot start a match is one step
        const uint32_t r7 = (e >> 8) & 0x7Fu, v = e & 0xFFu;
        return (!(e & 0x8000u) && r7 == 1u && v > 1u) ? (int)v : 1;
      };
      // first copy of the text at ta (length tl) in [from, ta), or ta
      auto index_of = [&](const int from, const int ta, const int tl) -> int {
        if (copies_are_matches && from >= P) {
          for (int u = (from - P) >> 6; u <= (ta - 1 - P) >> 6 && from < ta; u++) {
            unsigned long long mk = MM[u];
            const int ub0 = P + 64 * u;
            if (from > ub0) mk &= ~0ull << (from - ub0);
            if (ta < ub0 + 64) mk &= (1ull << (ta - ub0)) - 1ull;
            while (mk) {
              const int q = ub0 + __ffsll((long long)mk) - 1;
              mk &= mk - 1;
              if (same_text(chunk, q, ta, tl)) return q;
            }
          }
          return ta;
        }
        return (int)table_index_of_text(chunk, tab, from, ta, tl);
      };
      while (!stop && cu < lim) {
        const int tile_end = min(P + RT_TILE, dl);
        // 1. entries (positions outside [0, data_len) read as 0 and are never used)
#pragma unroll
        for (int q = 0; q < RT_TILE / 256; q++) {
          const int i0 = P + q * 256 + lane * 8;
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (i0 >= 0 && i0 < dl) v = *reinterpret_cast<const uint4*>(tab + i0);
          else if (i0 < 0 && i0 + 8 > 0) {
            uint16_t t8[8];
            for (int z = 0; z < 8; z++) t8[z] = i0 + z >= 0 ? tab[i0 + z] : (uint16_t)0;
            v = make_uint4(t8[0] | ((uint32_t)t8[1] << 16), t8[2] | ((uint32_t)t8[3] << 16), t8[4] | ((uint32_t)t8[5] << 16), t8[6] | ((uint32_t)t8[7] << 16));
          }
          *reinterpret_cast<uint4*>(E + q * 256 + lane * 8) = v;
        }
        __syncwarp();
        // 2. unit reach and match mask, cover entering the unit, start of the lane's chase
        const int ub = P + 64 * lane, ue = min(ub + 64, tile_end);
        int mreach = 0;
        unsigned long long mm = 0;
        for (int j = max(ub, 0); j < ue;) {
          const uint32_t e = E[j - P];
          mreach = max(mreach, j + (int)max((e >> 8) & 0x7Fu, e & 0xFFu));
          if (e & 0x8000u) mm |= 1ull << (j - ub);
          j += hop_of(e);
        }
        MM[lane] = mm;
        __syncwarp();
        int incl = mreach;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl = max(incl, y); }
        const int before = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
        const int cin = lane == 0 ? cover : max(cover, before);
        const int lane_cur = (cu - P) >> 6;
        int start = -1;
        if (lane == lane_cur) start = cu;
        else if (lane > lane_cur) {
          int lc = cin;
          for (int j = ub; j < ue && j < lim;) {
            if (lc <= j) { start = j; break; }
            const uint32_t e = E[j - P];
            lc = max(lc, j + (int)max((e >> 8) & 0x7Fu, e & 0xFFu));
            j += hop_of(e);
          }
        }
        const uint32_t startm = __ballot_sync(0xFFFFFFFFu, start >= 0);
        const uint32_t higher = lane == 31 ? 0u : (startm >> (lane + 1)) << (lane + 1);
        const int nstart = __shfl_sync(0xFFFFFFFFu, start, higher ? __ffs(higher) - 1 : lane);
        const int e_end = higher ? nstart : min(tile_end, lim);
        // 3. chase: matches and the last match end of every lane
        uint32_t n_l = 0;
        int last_mend = -1, a = start;
        if (start >= 0)
          while (a < e_end) {
            const uint32_t e = E[a - P];
            if (e & 0x8000u) { n_l++; last_mend = a + (int)(e & 0xFFu); }
            a += (int)(e & 0xFFu);
          }
        if (start >= 0 && higher && a != e_end) atomicOr(err, ERR_INTERNAL);   // (a sync point is visited by every replay)
        const int exit_a = a;
        int lm = last_mend;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xFFFFFFFFu, lm, o); if (lane >= o) lm = max(lm, y); }
        const int lm_before = __shfl_up_sync(0xFFFFFFFFu, lm, 1);
        const int sp_in = lane == 0 ? sp_w : max(sp_w, lm_before);
        const int lm_all = __shfl_sync(0xFFFFFFFFu, lm, 31);
        // 4. the reference's tests per match, in lane order
        uint32_t k_l = 0;
        int ev = 0;                       // 1: deferred (the chunk ends), 2: the text occurs earlier
        int ev_q = 0, ev_a = 0, ev_len = 0, ev_sp = 0;
        if (n_l) {
          int sp = sp_in;
          a = start;
          while (a < e_end) {
            const uint32_t e = E[a - P];
            const int v = (int)(e & 0xFFu);
            if (e & 0x8000u) {
              const int q = index_of(sp, a, v);
              if (full && q + v > defer_line) { ev = 1; break; }
              if (q < a) { ev = 2; ev_q = q; ev_a = a; ev_len = v; ev_sp = sp; break; }
              k_l++;
              sp = a + v;
            }
            a += v;
          }
        }
        const uint32_t evm = __ballot_sync(0xFFFFFFFFu, ev != 0);
        const int F = evm ? __ffs(evm) - 1 : 32;
        const uint32_t valid = lane < F ? n_l : lane == F ? k_l : 0u;
        uint32_t vincl = valid;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, vincl, o); if (lane >= o) vincl += y; }
        const uint32_t vtotal = __shfl_sync(0xFFFFFFFFu, vincl, 31);
        const int evF = __shfl_sync(0xFFFFFFFFu, ev, F & 31);
        const bool reloc = F < 32 && evF == 2;
        if (MODE == 2 && w + vtotal + (reloc ? 1u : 0u) > wend) { if (lane == 0) atomicOr(err, ERR_SLAB); n += vtotal + 1; stop = true; break; }
        if (MODE != 0 && valid) {
          unsigned long long wi = w + vincl - valid;
          int sp = sp_in;
          uint32_t left = valid;
          a = start;
          while (left) {
            const uint32_t e = E[a - P];
            const int v = (int)(e & 0xFFu);
            if (e & 0x8000u) {
              if (wi < wend) {
                ReaderHit h;
                h.search_abs = (long long)cstart + sp; h.d_true = (uint32_t)(a - sp); h.d_text = (uint32_t)(a - sp); h.chunk = k;
                hits[wi] = h;
              }
              wi++; left--;
              sp = a + v;
            }
            a += v;
          }
        }
        w += vtotal; n += vtotal;
        if (F == 32) {
          // the tile is done: on to the next one
          if (lm_all >= 0) sp_w = max(sp_w, lm_all);
          const int last = 31 - __clz((int)startm);
          cu = __shfl_sync(0xFFFFFFFFu, exit_a, last);
          cover = max(cover, __shfl_sync(0xFFFFFFFFu, incl, 31));
          P += RT_TILE;
        } else if (!reloc) {
          stop = true;     // too close to the boundary: the next chunk's job (and everything behind it)
        } else {
          const int q = __shfl_sync(0xFFFFFFFFu, ev_q, F), ta = __shfl_sync(0xFFFFFFFFu, ev_a, F);
          const int tl = __shfl_sync(0xFFFFFFFFu, ev_len, F), tsp = __shfl_sync(0xFFFFFFFFu, ev_sp, F);
          if (MODE != 0 && w < wend && lane == 0) {
            ReaderHit h;
            h.search_abs = (long long)cstart + tsp; h.d_true = (uint32_t)(ta - tsp); h.d_text = (uint32_t)(q - tsp); h.chunk = k;
            hits[w] = h;
          }
          w++; n++;
          sp_w = q + tl; cu = sp_w;
          P = P0 + ((cu - P0) & ~63);
          cover = cu + 254;
        }
        __syncwarp();
      }
      pos = sp_w; cur = cu;
      if (!stop && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(const_cast<int*>(slow_flag)) + 2, 1ull);   // statistics: chunks on the tile chase
    } else {
      if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(const_cast<int*>(slow_flag)) + 1, 1ull);   // statistics: chunks replayed sequentially
    }
    // the sequential form: the whole chunk, or what the tiles left (searchPos pos, next attempt at cur)
    bool resume = true;
    while (!stop && pos < data_len) {
      // FindBytesReuse(chunk[pos:data_len]): attempts at pos, f+1, ...
      int64_t a = resume ? max(cur, pos) : pos, mlen = -1;
      resume = false;
      int64_t wb = -64;
      uint32_t ew = 0;
      while (a < data_len) {
        if (!LINEAR && a >= wb + 32) {   // fetch the 32 entries from a on
          wb = a;
          ew = wb + lane < data_len ? tab[wb + lane] : 0u;
        }
        const uint32_t e = LINEAR ? 0u : __shfl_sync(0xFFFFFFFFu, ew, (int)(a - wb));
        const uint32_t r7 = (e >> 8) & 0x7Fu, v = e & 0xFFu;
        if (!LINEAR && r7 != RT_SLOW_REACH && v != RT_SLOW_VAL && a + (int64_t)r7 <= data_len) {   // (LINEAR: no table, every attempt runs the machine)
          if (e & 0x8000u) { mlen = (int64_t)v; break; }
          a += (int64_t)v;
        } else {
          int64_t ml = 0, ns = 0, reach = 0;
          if (reader_attempt(m, img, chunk, data_len, a, sc, err, &ml, &ns, &reach)) { mlen = ml; break; }
          if (ns > data_len) { a = data_len; break; }   // goto-machine: `if l > offset` fails, FindBytesReuse returns false
          a = ns;
        }
      }
      if (mlen < 0) break;
      const int64_t mstart = warp_index_of_text(chunk, pos, a, mlen, lane);
      const int64_t mend = mstart + mlen;
      if (full && mend > data_len - (int64_t)cp.L) break;  // too close to the boundary: next chunk's job
      if (MODE == 2 && w >= wend) { if (lane == 0) atomicOr(err, ERR_SLAB); break; }
      if (MODE != 0 && w < wend && lane == 0) {
        ReaderHit h;
        h.search_abs = (long long)cstart + pos; h.d_true = (uint32_t)(a - pos); h.d_text = (uint32_t)(mstart - pos); h.chunk = k;
        hits[w] = h;
      }
      w++; n++;
      if (mlen > 0) pos = mend; else pos++;
    }
    if (MODE != 1 && lane == 0) counts[j] = n;
  }
}

// Offset record (slice-relative) of the attempt at `start` of slice[0:sl), known to match.
__device__ __forceinline__ void reader_record(const DevMeta& m, const uint32_t* __restrict__ img, const uint8_t* __restrict__ slice,
                                              const int64_t sl, const int64_t start, const int nc, const Scratch& sc, int* err,
                                              int64_t* rec) {
  if (m.find_engine == FIND_TDFA) {
    int64_t mtags[MAX_CAPS];
    tdfa_walk(m, img, slice, sl, start, start == 0, mtags);
    rec[0] = mtags[0]; rec[1] = mtags[1];
    for (int g = 1; g < nc / 2; g++) {
      if (mtags[2 * g] >= 0) { rec[2 * g] = mtags[2 * g]; rec[2 * g + 1] = mtags[2 * g + 1]; }
      else { rec[2 * g] = -1; rec[2 * g + 1] = -1; }
    }
    return;
  }
  if (m.sl_n > 0 && m.sl_caps_ok) {
    // straight-line program: every group is set, at a fixed distance from the match start
    for (int g = 0; g < nc; g++) rec[g] = start + (int64_t)m.sl_cap[g];
    return;
  }
  int32_t caps[MAX_CAPS];
  bt_machine<MODE_FINDALL>(m, img, slice, sl, start, caps, sc, err);
  bt_emit_record(caps, nc, start, sl, 0, rec);
}

__global__ void __launch_bounds__(128) find_reader_records_kernel(const DevMeta m, const uint32_t* __restrict__ gimg, const int in_smem,
                                                                  const uint8_t* __restrict__ d_stream, const uint64_t base_off,
                                                                  const ChunkPlan cp, const ReaderHit* __restrict__ hits, const uint64_t n_hits,
                                                                  const uint64_t region, const uint64_t n_run,
                                                                  const unsigned long long* __restrict__ counts,
                                                                  const unsigned long long* __restrict__ bases,
                                                                  int64_t* __restrict__ out_soff, int32_t* __restrict__ out_chunk,
                                                                  int64_t* __restrict__ out_rec, const ScratchPlan sp, int* err) {
  extern __shared__ __align__(16) uint32_t smem_img[];
  __shared__ __align__(8) unsigned long long mbar;
  const uint32_t* img = gimg;
  if (in_smem) { stage_image_tma(smem_img, gimg, m.image_words, &mbar); img = smem_img; }
  const Scratch sc = scratch_of(sp);
  const int nc = m.find_engine == FIND_TDFA ? m.t_ntags : m.num_cap;
  // region == 0: hits[0:n_hits) is the compact list.  region > 0: chunk j's hits sit at hits[j*region ...) and go
  // to output positions bases[j] ...; a warp takes a chunk.
  if (region > 0) {
    const int lane = threadIdx.x & 31;
    for (uint64_t j = sc.tid >> 5; j < n_run; j += sp.stride >> 5) {
      const unsigned long long cnt = counts[j], ob = bases[j];
      for (unsigned long long t = lane; t < cnt; t += 32) {
        const unsigned long long i = ob + t;
        if (i >= n_hits) break;
        const ReaderHit h = hits[j * region + t];
        const uint64_t k = h.chunk;
        uint64_t cstart, dlen;
        bool full;
        chunk_geometry(cp, k, cstart, dlen, full);
        (void)full;
        const int64_t pos = h.search_abs - (long long)cstart;
        const uint8_t* slice = d_stream + ((uint64_t)h.search_abs - base_off);
        const int64_t sl = (int64_t)dlen - pos;
        int64_t rec[MAX_CAPS];
        reader_record(m, img, slice, sl, (int64_t)h.d_true, nc, sc, err, rec);
        out_soff[i] = h.search_abs + (long long)h.d_text;
        out_chunk[i] = (int32_t)k;
        for (int g = 0; g < nc; g++) out_rec[i * nc + g] = rec[g] < 0 ? -1 : h.search_abs + rec[g];
      }
    }
    return;
  }
  for (uint64_t i = sc.tid; i < n_hits; i += sp.stride) {
    const ReaderHit h = hits[i];
    const uint64_t k = h.chunk;
    uint64_t cstart, dlen;
    bool full;
    chunk_geometry(cp, k, cstart, dlen, full);
    (void)full;
    const int64_t pos = h.search_abs - (long long)cstart;
    const uint8_t* slice = d_stream + ((uint64_t)h.search_abs - base_off);
    const int64_t sl = (int64_t)dlen - pos;
    int64_t rec[MAX_CAPS];
    reader_record(m, img, slice, sl, (int64_t)h.d_true, nc, sc, err, rec);
    out_soff[i] = h.search_abs + (long long)h.d_text;
    out_chunk[i] = (int32_t)k;
    for (int g = 0; g < nc; g++) out_rec[i * nc + g] = rec[g] < 0 ? -1 : h.search_abs + rec[g];
  }
}

}  // namespace rgx
