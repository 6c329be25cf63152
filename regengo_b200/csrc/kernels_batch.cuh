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

// ---------------------------------------------------------------------------------------------------------------
// match_multi_kernel -- batched MatchBytes for MANY patterns in one launch (BASELINE.json configs[3]: the curated
// suite over 100 M short inputs).  The batch is packed bytes[] + offsets[n+1]; the inputs of program p are the index
// range [prog_first[p], prog_first[p+1]).  Work is cut into ITEMS = runs of consecutive 256-input tiles of one
// program; a CTA stages its program's image once (TMA bulk copy) and then every warp, 32 inputs at a time:
//   * reads the offsets with one coalesced load (32-bit shard-relative offsets on the batch path),
//   * copies the 32 inputs' bytes -- one contiguous span of the packed buffer -- into shared memory with 16-byte loads,
//   * lane t runs the pattern's engine on its input OUT OF SHARED MEMORY (no per-byte global loads, no sector
//     waste on 12-byte inputs) and writes its flag byte; a warp's 32 flags are one 32-byte store.
//   * LOCAL: the goto-machine's backtrack stack and visited bits live in per-thread LOCAL memory, which the L1 caches
//     write-back -- a push/pop costs an L1 hit instead of an L2 round trip to the interleaved global scratch.  An
//     input that outgrows the local arrays flags ERR_STACK / ERR_VISITED and the host re-runs the batch with the
//     global scratch sized for the provable worst case (LOCAL = false).
// One thread still runs one machine sequentially, so restart order and the skip-restart rule hold by
// construction.  A tile whose bytes do not fit the staging buffer (long inputs) reads global memory directly.
constexpr uint32_t MM_TILE = 256;                 // inputs per tile = threads per CTA
constexpr uint32_t MM_WARP_BYTES = 2048;          // per-warp staging buffer for the bytes of 32 inputs
constexpr uint32_t MM_TILE_BYTES = (MM_TILE / 32) * MM_WARP_BYTES;
constexpr uint32_t MM_IMAGE_BYTES = 40 * 1024;    // images up to this size are staged; larger ones are read from L2
constexpr uint32_t MM_LSTACK = 256;               // backtrack-stack entries kept in (L1-cached, write-back) local memory
constexpr uint32_t MM_LVISITED = 640;             // words of the memoisation bit-vector kept in local memory

static_assert(sizeof(DevMeta) % 4 == 0, "DevMeta is copied word by word");
struct MultiArgs {
  const DevMeta* metas;              // [n_progs]
  const uint32_t* const* images;     // [n_progs] device image words
  const uint32_t* item_base;         // [n_progs + 1] first item of each program
  const unsigned long long* prog_first;   // [n_progs + 1] first input of each program
  uint32_t n_progs, n_items, tiles_per_item;
  uint32_t image_bytes;              // shared memory reserved for a staged image (the largest one, at most MM_IMAGE_BYTES)
};

template <typename OFFT>
__global__ void max_len_multi_kernel(const OFFT* __restrict__ offs, uint64_t n, unsigned long long* out) {
  unsigned long long mx = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long d = (unsigned long long)(offs[i + 1] - offs[i]);
    mx = d > mx ? d : mx;
  }
  for (int o = 16; o; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xFFFFFFFFu, mx, o); mx = v > mx ? v : mx; }
  if ((threadIdx.x & 31) == 0 && mx) atomicMax(out, mx);
}

template <typename OFFT, bool LOCAL>
__global__ void __launch_bounds__(MM_TILE) match_multi_kernel(const MultiArgs a, const uint8_t* __restrict__ bytes,
                                                              const OFFT* __restrict__ offs, uint8_t* __restrict__ flag,
                                                              const ScratchPlan sp, int* err) {
  extern __shared__ __align__(16) uint32_t mm_smem[];       // [image | tile bytes]
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ DevMeta sm_meta;
  uint32_t* simg = mm_smem;
  uint8_t* tbytes = reinterpret_cast<uint8_t*>(mm_smem) + a.image_bytes;

  Scratch sc;
  sc.stack = sp.stack; sc.cstack = sp.cstack; sc.visited = sp.visited;
  sc.stack_cap = sp.stack_cap; sc.cstack_cap = sp.cstack_cap; sc.visited_words = sp.visited_words;
  sc.stride = sp.stride; sc.tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint2 lstack[LOCAL ? MM_LSTACK : 1];
  uint32_t lvisited[LOCAL ? MM_LVISITED : 1];
  if (LOCAL) {
    sc.stack = lstack; sc.visited = lvisited; sc.cstack = nullptr;
    sc.stack_cap = MM_LSTACK; sc.cstack_cap = 0; sc.visited_words = MM_LVISITED; sc.stride = 1; sc.tid = 0;
  }

  int cur_prog = -1;
  const uint32_t* img = nullptr;
  for (uint32_t item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    // which program: the last p with item_base[p] <= item
    uint32_t lo = 0, hi = a.n_progs;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.item_base[mid] <= item) lo = mid; else hi = mid; }
    const uint32_t p = lo;
    if ((int)p != cur_prog) {
      __syncthreads();   // everyone is done with the previous image / meta
      const uint32_t* src = reinterpret_cast<const uint32_t*>(a.metas + p);
      uint32_t* dst = reinterpret_cast<uint32_t*>(&sm_meta);
      for (uint32_t k = threadIdx.x; k < sizeof(DevMeta) / 4; k += MM_TILE) dst[k] = src[k];
      __syncthreads();
      img = a.images[p];
      if (sm_meta.match_words * 4u <= a.image_bytes) {
        stage_image_tma(simg, img, sm_meta.match_words, &mbar);
        __syncthreads();
        if (threadIdx.x == 0) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        img = simg;
      }
      cur_prog = (int)p;
    }
    const DevMeta& m = sm_meta;
    const unsigned long long in_first = a.prog_first[p], in_end = a.prog_first[p + 1];
    const unsigned long long item_first = in_first + (unsigned long long)(item - a.item_base[p]) * a.tiles_per_item * MM_TILE;
    const unsigned long long item_end = min(in_end, item_first + (unsigned long long)a.tiles_per_item * MM_TILE);
    // Each WARP walks the item's inputs 32 at a time on its own (no block barrier in this loop: a slow input holds
    // up its warp, not the other seven): offsets by one coalesced load, the 32 inputs' bytes -- one contiguous span
    // -- into the warp's staging buffer with 16-byte loads, one machine per lane out of shared memory.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* wbytes = tbytes + (size_t)warp * MM_WARP_BYTES;
    for (unsigned long long i0 = item_first + (unsigned long long)warp * 32; i0 < item_end; i0 += MM_TILE) {
      const uint32_t nin = (uint32_t)min(32ull, item_end - i0);
      const unsigned long long b = (uint32_t)lane < nin ? (unsigned long long)offs[i0 + lane] : 0ull;
      const unsigned long long e_last = (unsigned long long)offs[i0 + nin];
      const unsigned long long b_next = __shfl_down_sync(0xFFFFFFFFu, b, 1);
      const unsigned long long e = (uint32_t)lane + 1 < nin ? b_next : e_last;
      const unsigned long long b0 = __shfl_sync(0xFFFFFFFFu, b, 0), b1 = e_last;
      const uintptr_t addr0 = reinterpret_cast<uintptr_t>(bytes + b0) & ~(uintptr_t)15;
      const unsigned long long span = reinterpret_cast<uintptr_t>(bytes + b1) - addr0;   // bytes from the aligned base to the tile's end
      const bool staged = span + 16 <= MM_WARP_BYTES;
      __syncwarp();   // the previous tile's bytes are no longer read
      if (staged) {
        // (reads stay inside the 16-byte chunks that hold the tile's first and last byte)
        const uint4* g = reinterpret_cast<const uint4*>(addr0);
        uint4* d = reinterpret_cast<uint4*>(wbytes);
        const uint32_t nvec = b1 > b0 ? (uint32_t)((span - 1) >> 4) + 1u : 0u;
        for (uint32_t k = lane; k < nvec; k += 32) d[k] = g[k];
      }
      __syncwarp();
      if ((uint32_t)lane < nin) {
        const int64_t l = (int64_t)(e - b);
        const uint8_t* in = staged ? wbytes + (reinterpret_cast<uintptr_t>(bytes + b) - addr0) : bytes + b;
        int r;
        if (m.match_engine == MATCH_THOMPSON) r = thompson_match(m, img, in, l);
        else {
          if (LOCAL && (m.flags & F_MATCH_MEMO)) {
            // only the words this input needs are cleared per restart (compiler.go:812-818: numInst * (l+1) bits)
            const unsigned long long words = ((unsigned long long)m.n_inst * (unsigned long long)(l + 1) + 31) / 32;
            sc.visited_words = (uint32_t)min(words, (unsigned long long)MM_LVISITED);
            if (words > MM_LVISITED) atomicOr(err, ERR_VISITED);
          }
          r = bt_machine<MODE_MATCH>(m, img, in, l, 0, nullptr, sc, err);
        }
        flag[i0 + lane] = (uint8_t)r;
      }
    }
  }
}

}  // namespace rgx
