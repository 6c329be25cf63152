// findall_scan_linear_kernel -- the FindAll scan for a STRAIGHT-LINE backtracking program (device_program.cu: sl_*; no
// Alt, one ASCII class per step, captures at fixed distances: DateCapture, `(\d{4})-(\d{2})`, ...).
//
// The attempt started at s (find.go:130-316 tries every searchStart; no skip-restart in FindAll) matches exactly when
// the S bytes at s pass the S class tests, its length is S and its captures sit at fixed offsets -- so the records of
// a segment are the set bits of the `alive` plane of linear_unit_planes (kernels_stream.cuh), computed for 64 start
// positions per lane without running the goto-machine.  Same slab format as findall_scan_kernel<FIND_BT> (8 KiB
// segments); the cursor replay (chain) and the emit kernel are unchanged.
//
// FULL = false: the program only STARTS with a straight line (slp_n >= 2 steps before its first Alt / EmptyWidth /
// loop: TDFALogParser's 19-step timestamp).  Passing those steps is necessary for a match, so the plane is the
// candidate filter -- far sharper than the first-byte set -- and the goto-machine runs on its set bits only: once to
// count the matches of a lane (their slab slots are ordered), once more to write them.
#pragma once
#include "kernels_findall.cuh"
#include "kernels_stream.cuh"

namespace rgx {

template <bool FULL>
__global__ void __launch_bounds__(SCAN_WARPS * 32) findall_scan_linear_kernel(
    const DevMeta m, const uint32_t* __restrict__ gimg, const int in_smem, const uint8_t* __restrict__ buf, const uint64_t len,
    const uint32_t mis /* buf - align_down_16(buf) */, const uint64_t n_seg, const FindAllBufs fb, const ScratchPlan sp, int* err) {
  extern __shared__ __align__(16) uint32_t smem_all[];
  __shared__ __align__(8) unsigned long long mbar;
  const uint32_t* img = gimg;
  if (in_smem) { stage_image_tma(smem_all, gimg, m.image_words, &mbar); img = smem_all; }
  const uint8_t* cm = reinterpret_cast<const uint8_t*>(img + m.off_sl_cm);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint8_t* abuf = buf - mis;                       // 16-byte aligned view; valid bytes are [mis, mis + len)
  const uint64_t end_a = (uint64_t)mis + len;
  const int S = FULL ? m.sl_n : m.slp_n;
  const int nc = m.num_cap;
  const Scratch sc = scratch_of(sp);
  const uint64_t total_warps = (uint64_t)gridDim.x * SCAN_WARPS;

  for (uint64_t seg = (uint64_t)blockIdx.x * SCAN_WARPS + warp; seg < n_seg; seg += total_warps) {
    const uint64_t seg_a = seg * SEG_BYTES;              // aligned-space position of the segment
    uint32_t cnt = 0;
    for (uint32_t it = 0; it < SEG_BYTES / 2048; it++) {
      const uint32_t urel = it * 2048 + (uint32_t)lane * 64;
      const uint64_t ua = seg_a + urel;
      unsigned long long alive = 0;
      if (ua < end_a) {
        uint32_t bw[24];
        if (ua >= mis && ua + 96 <= end_a) {
#pragma unroll
          for (int q = 0; q < 6; q++) {
            const uint4 v = *reinterpret_cast<const uint4*>(abuf + ua + 16 * q);
            bw[4 * q] = v.x; bw[4 * q + 1] = v.y; bw[4 * q + 2] = v.z; bw[4 * q + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int q = 0; q < 24; q++) {
            uint32_t x = 0;
            for (int b = 0; b < 4; b++) {
              const uint64_t pp = ua + 4 * q + b;
              x |= (uint32_t)((pp >= mis && pp < end_a) ? abuf[pp] : 0xFFu) << (8 * b);   // bytes that do not exist are in no class
            }
            bw[q] = x;
          }
        }
        unsigned long long cand, pl[6];
        linear_unit_planes(m, S, cm, bw, cand, alive, pl);
        // starts outside [mis, mis + cand_len) do not belong to this buffer (a match needs its S bytes: bytes past the
        // end read as 0xFF and fail their step, so nothing has to be cut at the end)
        if (ua < mis) alive &= ~0ull << (mis - ua);
        const uint64_t cend = (uint64_t)mis + fb.cand_len;
        if (ua + 64 > cend) alive = ua >= cend ? 0ull : alive & ((1ull << (cend - ua)) - 1ull);
      }
      int32_t first_caps[MAX_CAPS];     // (!FULL) the result of the lane's first confirmed start: not computed twice
      if (!FULL) {
        // the plane is a filter here: keep the starts the goto-machine confirms
        unsigned long long rest = alive;
        bool have_first = false;
        while (rest) {
          const uint32_t j = (uint32_t)__ffsll((long long)rest) - 1;
          rest &= rest - 1;
          int32_t caprel[MAX_CAPS];
          if (!bt_machine<MODE_FINDALL>(m, img, buf, (int64_t)len, (int64_t)(ua + j) - (int64_t)mis, caprel, sc, err)) alive &= ~(1ull << j);
          else if (!have_first) {
            have_first = true;
            for (int g = 1; g < nc; g++) first_caps[g] = caprel[g];
          }
        }
      }
      bool first_record = true;
      const uint32_t c = (uint32_t)__popcll(alive);
      uint32_t incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += y; }
      const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
      uint32_t slot = cnt + incl - c;
      while (alive) {
        const uint32_t j = (uint32_t)__ffsll((long long)alive) - 1;
        alive &= alive - 1;
        if (slot < fb.K) {
          const uint64_t r = seg * fb.K + slot;
          if (FULL) {
            fb.keys[r] = make_uint2(urel + j, (uint32_t)S);
            for (int g = 2; g < nc; g++) fb.caps[r * fb.cw + (g - 2)] = (int32_t)m.sl_cap[g];
          } else if (first_record) {
            fb.keys[r] = make_uint2(urel + j, (uint32_t)first_caps[1]);
            for (int g = 2; g < nc; g++) fb.caps[r * fb.cw + (g - 2)] = first_caps[g];
          } else {
            int32_t caprel[MAX_CAPS];
            bt_machine<MODE_FINDALL>(m, img, buf, (int64_t)len, (int64_t)(ua + j) - (int64_t)mis, caprel, sc, err);
            fb.keys[r] = make_uint2(urel + j, (uint32_t)caprel[1]);
            for (int g = 2; g < nc; g++) fb.caps[r * fb.cw + (g - 2)] = caprel[g];
          }
        }
        first_record = false;
        slot++;
      }
      cnt += total;
    }
    if (cnt > fb.K && lane == 0) atomicOr(err, ERR_SLAB);
    if (lane == 0) fb.count[seg] = min(cnt, fb.K);
  }
}

}  // namespace rgx
