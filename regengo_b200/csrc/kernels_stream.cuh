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

}  // namespace rgx
