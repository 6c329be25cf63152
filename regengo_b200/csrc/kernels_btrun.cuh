// findall_scan_btrun_kernel -- FindAll scan for backtracking patterns of the shape
//
//        (cap|nop)*  C+  (cap|nop)*  b  rest          C an ASCII byte class, b a literal byte, b not in C
//
// e.g. (?P<user>\w+)@(?P<domain>\w+)\.(?P<tld>\w+).  Every \w would be a candidate start and every
// attempt would run to the end of its word before failing on '@'; instead the scan anchors on b.
//
// Why this is exact (generated machine: find.go:130-316, instructions.go:178-197, 331-457):
//   An attempt at s first runs the greedy loop C+ to the end p of the maximal C-run from s, then needs
//   input[p] == b.  Shorter loop counts end at a byte of C, which is not b, so backtracking into the
//   loop always fails.  Hence the attempt at s succeeds iff input[s] is in C, input[p] == b and `rest`
//   succeeds from p+1 -- and since `rest` runs from the same machine state for every s of the run
//   (captures written before the loop hold s itself, the ones between loop and b hold p), all starts
//   s0..p-1 of the maximal run before an occurrence of b share ONE outcome: same end, same captures
//   apart from the start-valued ones.  So one attempt at the run's first byte s0 decides them all, and
//   the result is stored as a RUN RECORD {first = s0, last = p-1, end, captures relative to s0}.  The
//   chain (kernels_chain.cuh, FIND_BT_RUN) then takes max(cursor, first) as the reference does when its
//   cursor lands inside a run, and the emit kernel substitutes that start into the start-valued captures.
//
// One warp per 32 KiB segment; a record belongs to the segment that holds its b.
#pragma once
#include "kernels_findall2.cuh"

namespace rgx {

constexpr uint32_t SEGB_BYTES = 32768;   // ~80 candidates per segment on log text: 2-3 nearly full verify batches
constexpr uint32_t QBCAP = 1024;
constexpr int BTRUN_WARPS = 8;

__global__ void __launch_bounds__(BTRUN_WARPS * 32, 4) findall_scan_btrun_kernel(
    const DevMeta m, const uint32_t* __restrict__ gimg, const int in_smem, const uint8_t* __restrict__ buf, const uint64_t len,
    const uint32_t mis, const uint64_t n_seg, const FindAllBufs fb, const ScratchPlan sp, int* err) {
  extern __shared__ __align__(16) uint32_t smem_all[];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint16_t queue[BTRUN_WARPS][QBCAP];
  const uint32_t* img = gimg;
  if (in_smem) { stage_image_tma(smem_all, gimg, m.image_words, &mbar); img = smem_all; }

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint16_t* q = queue[warp];
  Scratch sc;
  sc.stack = sp.stack; sc.cstack = sp.cstack; sc.visited = sp.visited;
  sc.stack_cap = sp.stack_cap; sc.cstack_cap = sp.cstack_cap; sc.visited_words = sp.visited_words;
  sc.stride = sp.stride; sc.tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint8_t* abuf = buf - mis;
  const uint64_t end_a = (uint64_t)mis + len;
  const uint32_t pb = (uint32_t)m.run_lit * 0x01010101u;
  const uint32_t* cls = img + m.off_cls + 8 * m.run_class_pc;
  const uint64_t total_warps = (uint64_t)gridDim.x * BTRUN_WARPS;
  const int nc = m.num_cap;

  for (uint64_t seg = (uint64_t)blockIdx.x * BTRUN_WARPS + warp; seg < n_seg; seg += total_warps) {
    const uint64_t seg_a = seg * SEGB_BYTES;
    const bool interior = seg_a >= mis && seg_a + SEGB_BYTES <= end_a;
    uint32_t tail = 0;
    bool dense = false;
    // FILTER: positions p with input[p] == b.  A lane owns 64 contiguous bytes of each 2 KiB block; SWAR
    // zero-byte test on w ^ b (the cheap test can only err on the byte right above a true hit, i.e. when two
    // hits touch; that redoes the block with the exact test), flags packed to two words per lane (bit
    // 8*byte + word), ordered append through a warp prefix sum.  Whether a class run ends at p is left to
    // the verify step (an empty run is no candidate).
    const bool interior_ld = interior;
    auto load_guarded = [&](const uint64_t apos) -> uint4 {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (apos + 16 > mis && apos < end_a) {
        uint8_t* vb = reinterpret_cast<uint8_t*>(&v);
        for (int j = 0; j < 16; j++) if (apos + j >= mis && apos + j < end_a) vb[j] = abuf[apos + j];
      }
      return v;
    };
    uint32_t w[16];
    auto load_block = [&](const uint32_t blk) {
      const uint64_t off = seg_a + (uint64_t)blk * 2048 + (uint64_t)lane * 64;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint64_t apos = off + u * 16;
        const uint4 v = (interior_ld || (apos >= mis && apos + 16 <= end_a)) ? *reinterpret_cast<const uint4*>(abuf + apos) : load_guarded(apos);
        w[4 * u] = v.x; w[4 * u + 1] = v.y; w[4 * u + 2] = v.z; w[4 * u + 3] = v.w;
      }
    };
    constexpr uint32_t N_BLK = SEGB_BYTES / 2048;
    const bool zero_lit = m.run_lit == 0;   // out-of-range bytes read as 0: clip even interior-looking hits then
    load_block(0);
    for (uint32_t blk = 0; blk < N_BLK; blk++) {
      uint32_t e0 = 0, e1 = 0;
#pragma unroll
      for (int j = 0; j < 16; j++) {
        const uint32_t z = w[j] ^ pb;
        const uint32_t d = (z - 0x01010101u) & ~z & 0x80808080u;
        if (j < 8) e0 |= d >> (7 - j); else e1 |= d >> (15 - j);
      }
      if ((e0 & (e0 >> 8)) | (e1 & (e1 >> 8))) {
        e0 = 0; e1 = 0;
#pragma unroll
        for (int j = 0; j < 16; j++) {
          const uint32_t z = w[j] ^ pb;
          const uint32_t d = ~(((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z | 0x7F7F7F7Fu);
          if (j < 8) e0 |= d >> (7 - j); else e1 |= d >> (15 - j);
        }
      }
      if (blk + 1 < N_BLK) load_block(blk + 1);
      if ((!interior || zero_lit) && (e0 | e1)) {
        const uint64_t apos = seg_a + (uint64_t)blk * 2048 + (uint64_t)lane * 64;
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
          const uint32_t wb = tail + incl - c;
          uint32_t pbase = blk * 2048 + lane * 64;
          uint32_t i = 0, e = e0;
          for (;;) {
            if (!e) { if (!e1) break; e = e1; e1 = 0; pbase += 32; }
            const uint32_t b = __ffs(e) - 1;
            e &= e - 1;
            const uint32_t pos = pbase + 4 * (b & 7u) + (b >> 3);
            uint32_t j = i;
            while (j > 0 && wb + j - 1 < QBCAP && q[wb + j - 1] > pos) { if (wb + j < QBCAP) q[wb + j] = q[wb + j - 1]; j--; }
            if (wb + j < QBCAP) q[wb + j] = (uint16_t)pos;
            i++;
          }
        }
        tail += total;
        if (tail > QBCAP) { dense = true; tail = QBCAP; }
      }
    }
    __syncwarp();
    if (dense && lane == 0) atomicOr(err, ERR_DENSE);

    // VERIFY: walk back to the start of the C-run, one exact attempt of the goto-machine from there
    const uint32_t n = tail;
    for (uint32_t base = 0; base < n; base += 32) {
      const uint32_t k = base + lane;
      if (k < n) {
        const int64_t p = (int64_t)(seg_a + q[k]) - (int64_t)mis;   // buffer-relative position of b
        int64_t s0 = p;
        while (s0 > 0) {
          const uint32_t pc = buf[s0 - 1];
          if (!((cls[pc >> 5] >> (pc & 31)) & 1u)) break;
          s0--;
        }
        int32_t caps[MAX_CAPS];
        const uint64_t r = seg * fb.K + k;
        uint2 key = make_uint2(0, KEY_INVALID);
        bool matched;
        if (s0 == p) {
          matched = false;   // no class byte before b: C+ cannot match here
        } else if (m.lin_n > 0) {
          // straight-line continuation (device_program.cu): the greedy path is the only one that can succeed
          for (int i = 0; i < nc; i++) caps[i] = ((m.run_start_caps >> i) & 1u) ? 0 : CAP_ZERO;
          int64_t pos = p;
          matched = false;
          bool alive = true;
          for (int e = 0; e < m.lin_n && alive; e++) {
            const uint32_t el = m.lin[e], arg = el >> 8;
            switch (el & 255u) {
              case LIN_CAP: caps[arg] = (int32_t)(pos - s0); break;
              case LIN_LIT:
                if (pos >= (int64_t)len || buf[pos] != (uint8_t)arg) alive = false; else pos++;
                break;
              case LIN_CLS: {
                const uint32_t* bm = img + m.off_cls + 8 * arg;
                const uint32_t c = pos < (int64_t)len ? buf[pos] : 256u;
                if (c > 255u || !((bm[c >> 5] >> (c & 31)) & 1u)) alive = false; else pos++;
                break;
              }
              case LIN_LOOP: {
                const uint32_t* bm = img + m.off_cls + 8 * arg;
                while (pos < (int64_t)len) {
                  const uint32_t c = buf[pos];
                  if (!((bm[c >> 5] >> (c & 31)) & 1u)) break;
                  pos++;
                }
                break;
              }
              default:   // LIN_MATCH
                caps[1] = (int32_t)(pos - s0);
                matched = true;
                alive = false;
                break;
            }
          }
          if (pos - s0 > 0x7FFFFFFFll) { atomicOr(err, ERR_RANGE); matched = false; }
        } else {
          matched = bt_machine<MODE_FINDALL>(m, img, buf, (int64_t)len, s0, caps, sc, err, m.run_resume_pc, p, m.run_start_caps) != 0;
        }
        if (matched) {
          const int64_t first_rel = s0 - ((int64_t)seg_a - (int64_t)mis);   // may be negative: the run began in an earlier segment
          const int64_t run_extra = p - 1 - s0;
          const int64_t mlen = caps[1];
          if (run_extra > 0xFFF || mlen >= (1 << 20) || first_rel < INT32_MIN) atomicOr(err, ERR_DENSE);  // generic scan instead
          key = make_uint2((uint32_t)(int32_t)first_rel, ((uint32_t)mlen << 12) | (uint32_t)run_extra);
          if (k < fb.K) for (int j = 2; j < nc; j++) fb.caps[r * fb.cw + (j - 2)] = caps[j];
        }
        if (k < fb.K) fb.keys[r] = key; else atomicOr(err, ERR_SLAB);
      }
      __syncwarp();
    }
    if (lane == 0) fb.count[seg] = min(n, fb.K);
  }
}

}  // namespace rgx
