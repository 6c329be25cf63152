// Batched MatchBytes / FindBytes: one thread per input runs the pattern's engine exactly as the
// generated method would (thread-per-machine keeps the reference's restart order).
// Inputs are packed bytes[] + offsets[n+1]; consecutive threads own consecutive inputs, so a warp's
// reads fall in one or two neighbouring 128-byte lines.
#pragma once
#include "engines.cuh"

namespace rgx {

struct ScratchPlan {
  uint2* stack; int32_t* cstack; uint32_t* visited;
  uint32_t stack_cap, cstack_cap, visited_words, stride;
};

__global__ void max_len_kernel(const uint64_t* __restrict__ offs, uint64_t n, unsigned long long* out) {
  unsigned long long mx = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long d = offs[i + 1] - offs[i];
    mx = d > mx ? d : mx;
  }
  for (int o = 16; o; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xFFFFFFFFu, mx, o); mx = v > mx ? v : mx; }
  if ((threadIdx.x & 31) == 0) atomicMax(out, mx);
}

// WHAT: 0 = MatchBytes -> flag[i]; 1 = FindBytes -> flag[i], rec[i*num_cap..]
template <int WHAT>
__global__ void __launch_bounds__(256) batch_kernel(const DevMeta m, const uint32_t* __restrict__ gimg, const int in_smem,
                                                    const uint8_t* __restrict__ bytes, const uint64_t* __restrict__ offs,
                                                    const uint64_t n, uint8_t* __restrict__ flag, int64_t* __restrict__ rec,
                                                    const ScratchPlan sp, int* err) {
  extern __shared__ __align__(16) uint32_t smem_img[];
  __shared__ __align__(8) unsigned long long mbar;
  const uint32_t* img = gimg;
  if (in_smem) { stage_image_tma(smem_img, gimg, m.image_words, &mbar); img = smem_img; }

  Scratch sc;
  sc.stack = sp.stack; sc.cstack = sp.cstack; sc.visited = sp.visited;
  sc.stack_cap = sp.stack_cap; sc.cstack_cap = sp.cstack_cap; sc.visited_words = sp.visited_words;
  sc.stride = sp.stride; sc.tid = blockIdx.x * blockDim.x + threadIdx.x;

  for (uint64_t i = sc.tid; i < n; i += sp.stride) {
    const uint64_t b = offs[i];
    const int64_t l = (int64_t)(offs[i + 1] - b);
    const uint8_t* in = bytes + b;
    if (WHAT == 0) {
      int r;
      if (m.match_engine == MATCH_THOMPSON) r = thompson_match(m, img, in, l);
      else r = bt_machine<MODE_MATCH>(m, img, in, l, 0, nullptr, sc, err);
      flag[i] = (uint8_t)r;
    } else {
      int64_t* out = rec + i * (uint64_t)m.num_cap;
      int r;
      if (m.find_engine == FIND_TDFA) {
        r = tdfa_find(m, img, in, l, 0, out);
      } else {
        int32_t caps[MAX_CAPS];
        r = bt_machine<MODE_FIND>(m, img, in, l, 0, caps, sc, err);
        if (r) bt_emit_record(caps, m.num_cap, 0, l, 0, out);
      }
      flag[i] = (uint8_t)r;
    }
  }
}

}  // namespace rgx
