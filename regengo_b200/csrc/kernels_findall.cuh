// FindAllBytes over one large device-resident buffer (find.go:130-316 for the backtracking
// engine, compiler.go:602-655 + tdfa.go:831-1052 for the TDFA engine), split into three phases
// that together reproduce the reference's sequential iteration bit for bit:
//
//  1. scan    (HBM-bound, the dominant kernel).  The buffer is cut into SEG-byte segments, one
//             warp per segment.  Lanes stream the segment with 16-byte coalesced loads, build a
//             candidate-start mask per 16 bytes (literal-prefix byte compare / first-byte set),
//             compact the candidates in order with a warp prefix sum, and verify 32 candidates at
//             a time with the pattern's exact engine started at that position (one TDFA walk or
//             one goto-machine attempt per lane).  Every start that matches becomes a RECORD
//             (start, length, capture offsets) in the segment's slab -- the outcome of "an
//             attempt at s" does not depend on where the reference's cursor was.
//  2. chain   (tiny).  The reference's cursor rule is replayed over the sparse records:
//             backtracking FindAll keeps the first record at/after the cursor and jumps to its end;
//             the TDFA FindAll advances the cursor by the match LENGTH from the slice start
//             (SURVEY Q2), so a record is returned floor(gap/len)+1 times.  Parts of G segments
//             are replayed independently from a speculated entry cursor and re-run until every
//             part's entry equals its predecessor's exit -- then the result is the sequential one.
//  3. emit    compacts the kept records, in order, into int64 offset records + repeat counts.
#pragma once
#include "engines.cuh"
#include "kernels_batch.cuh"

namespace rgx {

constexpr uint32_t SEG_BYTES = 8192;      // bytes per segment (one warp)
constexpr uint32_t QCAP = 1024;           // per-warp candidate ring (u16 segment-relative positions)
constexpr int SCAN_WARPS = 8;

struct FindAllBufs {
  uint32_t* count;     // [n_seg] records in the segment's slab
  uint2* keys;         // [n_seg * K] {start_rel, len}
  int32_t* caps;       // [n_seg * K * cw] capture offsets relative to the match start
  uint32_t* reps;      // [n_seg * K] chain output: times returned (0 = skipped)
  uint32_t K, cw;      // slab capacity per segment, ints per record in caps
  // sharding of one logical buffer over GPUs (defaults: cand_len = len, entry0 = 0, not_last = 0, out_base = 0)
  uint64_t cand_len;   // only starts < cand_len belong to this shard (the rest of the buffer is halo)
  long long entry0;    // cursor entering the shard (shard-relative)
  int not_last;        // the buffer end is not the end of the logical input
  long long out_base;  // added to every offset written by the emit kernel
};

// exact per-byte equality flags: 0x80 in every byte of x that equals the corresponding byte of pat
__device__ __forceinline__ uint32_t eq_bytes(uint32_t w, uint32_t pat) {
  const uint32_t x = w ^ pat;
  uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
  t = ~(t | x | 0x7F7F7F7Fu);
  return t;
}
__device__ __forceinline__ uint32_t gather4(uint32_t t) {  // flags at bits 7/15/23/31 -> low nibble
  return (((t >> 7) * 0x00204081u) >> 21) & 0xFu;
}

// ENGINE: FIND_BT or FIND_TDFA
template <int ENGINE>
__global__ void __launch_bounds__(SCAN_WARPS * 32) findall_scan_kernel(
    const DevMeta m, const uint32_t* __restrict__ gimg, const int in_smem, const uint8_t* __restrict__ buf, const uint64_t len,
    const uint32_t mis /* buf - align_down_16(buf) */, const uint64_t n_seg, const FindAllBufs fb, const ScratchPlan sp, int* err) {
  extern __shared__ __align__(16) uint32_t smem_all[];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint16_t queue[SCAN_WARPS][QCAP];
  const uint32_t* img = gimg;
  if (in_smem) { stage_image_tma(smem_all, gimg, m.image_words, &mbar); img = smem_all; }

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint16_t* q = queue[warp];
  Scratch sc;
  sc.stack = sp.stack; sc.cstack = sp.cstack; sc.visited = sp.visited;
  sc.stack_cap = sp.stack_cap; sc.cstack_cap = sp.cstack_cap; sc.visited_words = sp.visited_words;
  sc.stride = sp.stride; sc.tid = blockIdx.x * blockDim.x + threadIdx.x;

  const uint8_t* abuf = buf - mis;          // 16-byte aligned view; valid bytes are [mis, mis+len)
  // candidate starts are s in [0, len) -- plus s == len for a nullable TDFA (empty match at EOT)
  const uint64_t cand_end = mis + len + ((ENGINE == FIND_TDFA && m.nullable) ? 1 : 0);
  const uint32_t first_w0 = img[m.off_first + 0];
  const uint32_t pat = (uint32_t)m.prefix_bytes[0] * 0x01010101u;
  const uint64_t total_warps = (uint64_t)gridDim.x * SCAN_WARPS;

  for (uint64_t seg = (uint64_t)blockIdx.x * SCAN_WARPS + warp; seg < n_seg; seg += total_warps) {
    const uint64_t seg_a = seg * SEG_BYTES;  // aligned-space position of the segment
    uint32_t cnt = 0, head = 0, tail = 0;
    const uint32_t n_it = SEG_BYTES / 512;
    for (uint32_t it = 0; it <= n_it; it++) {
      if (it < n_it) {
        const uint64_t apos = seg_a + (uint64_t)it * 512 + (uint64_t)lane * 16;
        uint32_t mask = 0;
        if (apos < cand_end) {
          uint4 v = make_uint4(0, 0, 0, 0);
          if (apos >= mis && apos + 16 <= mis + len) {
            v = *reinterpret_cast<const uint4*>(abuf + apos);
          } else {
            uint8_t* vb = reinterpret_cast<uint8_t*>(&v);
            for (int j = 0; j < 16; j++) if (apos + j >= mis && apos + j < mis + len) vb[j] = abuf[apos + j];
          }
          if (m.gen_kind == GEN_PREFIX) {
            mask = gather4(eq_bytes(v.x, pat)) | (gather4(eq_bytes(v.y, pat)) << 4) | (gather4(eq_bytes(v.z, pat)) << 8) |
                   (gather4(eq_bytes(v.w, pat)) << 12);
          } else if (m.gen_kind == GEN_BYTESET) {
            const uint32_t* fs = img + m.off_first;
            const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 16; j++) {
              const uint32_t c = (ws[j >> 2] >> ((j & 3) * 8)) & 255u;
              mask |= ((fs[c >> 5] >> (c & 31)) & 1u) << j;
            }
          } else {
            mask = 0xFFFFu;
          }
          // drop positions outside [mis, cand_end)
          if (apos < mis) mask &= ~((1u << (uint32_t)(mis - apos)) - 1u);
          if (apos + 16 > cand_end) mask &= (1u << (uint32_t)(cand_end - apos)) - 1u;
          // verify the rest of a literal prefix right here (hits are rare)
          if (m.gen_kind == GEN_PREFIX && m.prefix_len > 1) {
            uint32_t mm = mask;
            while (mm) {
              const int j = __ffs(mm) - 1;
              mm &= mm - 1;
              const uint64_t p = apos + j;
              bool ok = p + m.prefix_len <= mis + len;
              for (int k = 1; ok && k < m.prefix_len; k++) ok = abuf[p + k] == m.prefix_bytes[k];
              if (!ok) mask &= ~(1u << j);
            }
          }
        }
        (void)first_w0;
        // ordered enqueue: warp exclusive prefix sum of the per-lane candidate counts
        const uint32_t c = __popc(mask);
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += y; }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        uint32_t w = tail + incl - c;
        uint32_t mm = mask;
        while (mm) {
          const int j = __ffs(mm) - 1;
          mm &= mm - 1;
          q[w & (QCAP - 1)] = (uint16_t)(it * 512 + lane * 16 + j);
          w++;
        }
        tail += total;
        __syncwarp();
      }
      // verify queued candidates, 32 at a time, in position order
      while (tail - head >= 32 || (it == n_it && tail != head)) {
        const uint32_t nb = min(tail - head, 32u);
        bool matched = false;
        uint32_t srel = 0, mlen = 0;
        int32_t caprel[MAX_CAPS];
        if (lane < (int)nb) {
          srel = q[(head + lane) & (QCAP - 1)];
          const int64_t s = (int64_t)(seg_a + srel) - (int64_t)mis;  // buffer-relative start
          if (ENGINE == FIND_TDFA) {
            int64_t mt[MAX_CAPS];
            const int64_t me = tdfa_walk(m, img, buf, (int64_t)len, s, false, mt);
            if (me >= 0) {
              matched = true;
              mlen = (uint32_t)(me - s);
              if (me - s > 0x7FFFFFFFll) atomicOr(err, ERR_RANGE);
              for (int j = 2; j < m.t_ntags; j++) caprel[j] = mt[j] < 0 ? -1 : (int32_t)(mt[j] - s);
            }
          } else {
            if (bt_machine<MODE_FINDALL>(m, img, buf, (int64_t)len, s, caprel, sc, err)) {
              matched = true;
              mlen = (uint32_t)caprel[1];
            }
          }
        }
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, matched);
        if (matched) {
          const uint32_t slot = cnt + __popc(bal & ((1u << lane) - 1u));
          if (slot < fb.K) {
            const uint64_t r = seg * fb.K + slot;
            fb.keys[r] = make_uint2(srel, mlen);
            const int nc = ENGINE == FIND_TDFA ? m.t_ntags : m.num_cap;
            for (int j = 2; j < nc; j++) fb.caps[r * fb.cw + (j - 2)] = caprel[j];
          } else {
            atomicOr(err, ERR_SLAB);
          }
        }
        cnt += __popc(bal);
        head += nb;
        __syncwarp();
      }
    }
    if (lane == 0) fb.count[seg] = min(cnt, fb.K);
  }
}

// ---- phase 2: replay the reference's cursor rule over the records ----------------------------------
struct ChainBufs {
  long long* exit_prev;   // [n_parts] exit cursor of each part in the previous pass
  long long* exit_cur;    // [n_parts]
  unsigned long long* part_sel;   // [n_parts] records kept in the part
  unsigned long long* part_reps;  // [n_parts] sum of reps in the part
  int* changed;
};

// ---- generic exact fallback: the reference loop, literally, on ONE device thread ---------------------
// Used for patterns the parallel path does not cover (anchored, memoised FindAll whose visited bits
// persist across iterations -- SURVEY Q12, TDFA whose begin/any start states differ -- Q3).
__global__ void findall_sequential_kernel(const DevMeta m, const uint32_t* __restrict__ img, const uint8_t* __restrict__ buf,
                                          const uint64_t len, const long long n_limit, int64_t* __restrict__ out,
                                          uint32_t* __restrict__ out_reps, const uint64_t cap_records, const ScratchPlan sp,
                                          unsigned long long* totals, int* err) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Scratch sc;
  sc.stack = sp.stack; sc.cstack = sp.cstack; sc.visited = sp.visited;
  sc.stack_cap = sp.stack_cap; sc.cstack_cap = sp.cstack_cap; sc.visited_words = sp.visited_words;
  sc.stride = 1; sc.tid = 0;
  const int64_t l = (int64_t)len;
  unsigned long long count = 0;
  if (n_limit != 0) {
    if (m.find_engine == FIND_TDFA) {
      const int nc = m.t_ntags;
      int64_t offset = 0, rec[MAX_CAPS];
      while (offset < l) {
        if (!tdfa_find(m, img, buf + offset, l - offset, offset, rec)) break;
        if (count < cap_records) { for (int j = 0; j < nc; j++) out[count * nc + j] = rec[j]; out_reps[count] = 1; }
        count++;
        if (n_limit > 0 && count >= (unsigned long long)n_limit) break;
        const int64_t ml = rec[1] - rec[0];
        offset += ml > 0 ? ml : 1;
      }
    } else {
      const int nc = m.num_cap;
      int32_t caps[MAX_CAPS];
      int64_t ss = 0;
      if (m.flags & F_FIND_MEMO) clear_visited(sc);
      for (;;) {
        if (n_limit > 0 && count >= (unsigned long long)n_limit) break;
        if ((m.flags & F_ANCHORED) && ss > 0) break;
        if (ss >= l) break;
        if (bt_machine<MODE_FINDALL>(m, img, buf, l, ss, caps, sc, err)) {
          if (count < cap_records) { bt_emit_record(caps, nc, ss, l, 0, out + count * nc); out_reps[count] = 1; }
          count++;
          ss = caps[1] > 0 ? ss + caps[1] : ss + 1;
        } else {
          ss++;
        }
      }
    }
  }
  totals[0] = count;
  totals[1] = count;
}

}  // namespace rgx
