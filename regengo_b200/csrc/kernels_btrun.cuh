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

__global__ void __launch_bounds__(BTRUN_WARPS * 32) findall_scan_btrun_kernel(
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
  const uint32_t lt_mask = (1u << lane) - 1u;
  constexpr uint32_t N_IT = SEGB_BYTES / 512;
  constexpr int U = 4;
  const int nc = m.num_cap;

  for (uint64_t seg = (uint64_t)blockIdx.x * BTRUN_WARPS + warp; seg < n_seg; seg += total_warps) {
    const uint64_t seg_a = seg * SEGB_BYTES;
    const bool interior = seg_a >= mis && seg_a + SEGB_BYTES <= end_a;
    uint32_t tail = 0;
    bool dense = false;
    auto load16 = [&](uint32_t it) -> uint4 {
      const uint64_t apos = seg_a + (uint64_t)it * 512 + (uint64_t)lane * 16;
      if (interior || (apos >= mis && apos + 16 <= end_a)) return *reinterpret_cast<const uint4*>(abuf + apos);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (apos + 16 > mis && apos < end_a) {
        uint8_t* vb = reinterpret_cast<uint8_t*>(&v);
        for (int j = 0; j < 16; j++) if (apos + j >= mis && apos + j < end_a) vb[j] = abuf[apos + j];
      }
      return v;
    };
    // FILTER: positions p with input[p] == b (approximate compare; verified below)
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
        cc[u][0] = eq_approx(cur[u].x, pb); cc[u][1] = eq_approx(cur[u].y, pb);
        cc[u][2] = eq_approx(cur[u].z, pb); cc[u][3] = eq_approx(cur[u].w, pb);
        any_all |= cc[u][0] | cc[u][1] | cc[u][2] | cc[u][3];
      }
      if (__ballot_sync(0xFFFFFFFFu, any_all != 0)) {
#pragma unroll
        for (int u = 0; u < U; u++) {
          uint32_t mask = 0;
          if (cc[u][0] | cc[u][1] | cc[u][2] | cc[u][3]) {
            mask = gather4(cc[u][0]) | (gather4(cc[u][1]) << 4) | (gather4(cc[u][2]) << 8) | (gather4(cc[u][3]) << 12);
            const uint64_t apos = seg_a + (uint64_t)(it0 + u) * 512 + (uint64_t)lane * 16;
            // exact test of b, and the byte before it must belong to C (else no run ends here)
            uint32_t mm = mask;
            while (mm) {
              const int j = __ffs(mm) - 1;
              mm &= mm - 1;
              const uint64_t ap = apos + j;
              bool ok = ap > mis && ap < end_a && abuf[ap] == (uint8_t)m.run_lit;
              if (ok) { const uint32_t pc = abuf[ap - 1]; ok = (cls[pc >> 5] >> (pc & 31)) & 1u; }
              if (!ok) mask &= ~(1u << j);
            }
          }
          const uint32_t bal = __ballot_sync(0xFFFFFFFFu, mask != 0);
          if (bal) {
            const uint32_t multi = __ballot_sync(0xFFFFFFFFu, (mask & (mask - 1)) != 0);
            if (!multi) {
              const uint32_t slot = tail + __popc(bal & lt_mask);
              if (mask && slot < QBCAP) q[slot] = (uint16_t)((it0 + u) * 512 + lane * 16 + __ffs(mask) - 1);
              tail += __popc(bal);
            } else {
              const uint32_t c = __popc(mask);
              uint32_t incl = c;
#pragma unroll
              for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += y; }
              uint32_t w = tail + incl - c;
              while (mask) {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1;
                if (w < QBCAP) q[w] = (uint16_t)((it0 + u) * 512 + lane * 16 + j);
                w++;
              }
              tail += __shfl_sync(0xFFFFFFFFu, incl, 31);
            }
            if (tail > QBCAP) dense = true;
          }
        }
      }
    }
    __syncwarp();
    if (dense) { if (lane == 0) atomicOr(err, ERR_DENSE); tail = QBCAP; }

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
        if (m.lin_n > 0) {
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
