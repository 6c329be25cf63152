// Device half of the C ABI (include/regengo_b200.h): contexts, per-device program images and the
// launch logic of every compute entry point.  There is no CPU matching path in this library: if
// CUDA is unavailable every function here fails with RGX_ECUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "capi_internal.hpp"
#include "device_program.cuh"
#include "engines.cuh"
#include "kernels_batch.cuh"
#include "kernels_findall.cuh"
#include "kernels_findall2.cuh"
#include "kernels_chain.cuh"
#include "kernels_scan6.cuh"
#include "kernels_emit.cuh"
#include "kernels_btrun.cuh"
#include "kernels_stream.cuh"
#include "kernels_replace.cuh"
#include "kernels_scanlin.cuh"
#include "replace_template.hpp"

using namespace rgx;

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

}  // namespace

struct rgx_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr, d2h_stream = nullptr;   // host-buffer FindAll: input upload / result download overlap the kernels
  uint64_t chunk_bytes = 256ull << 20;                         // upload granularity of the pipelined host-buffer FindAll
  std::vector<cudaEvent_t> chunk_ev;                           // one "uploaded" event per input chunk
  int sm_count = 0;
  int64_t launches = 0;
  // grow-only scratch
  DevBuf stack, cstack, visited, small, in_bytes, in_offs, out_flag, out_rec, out_reps, out_aux;
  DevBuf rp_tmpl, rp_len, rp_offs, rp_out;   // replace batch: template image, per-input lengths, host-call output staging
  DevBuf mm_tab;                       // match_multi: metas | image pointers | item_base | prog_first
  std::vector<uint8_t> mm_key;         // host copy of the last table uploaded (skips the upload when nothing changed)
  DevBuf fa_count, fa_keys, fa_caps, fa_reps, ch_a, ch_b, ch_sel, ch_reps, ch_selbase, ch_repsbase, ch_segsel, ch_segreps, ch_entry, ch_tile;
  void* h_small = nullptr;  // pinned, 4 KiB
  uint32_t fa_K = 128;      // slab capacity per segment, doubled on overflow
  uint32_t fa_stack_cap = 256;
  int64_t stat_seq_chunks = 0;   // FindReader: chunks of the last call that took the sequential replay (statistics)
  int chain_first_batch = 8;  // chain passes queued before the first readback (adapts to the data)
  // optional per-phase timing of the last FindAll (CUDA events on the launching stream)
  bool timing = false;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  float last_ms[4] = {0, 0, 0, 0};  // scan, chain, emit, total
};

namespace {

#define CU(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                                 \
      return RGX_ECUDA;                                                                              \
    }                                                                                                \
  } while (0)

int ensure(rgx_ctx* c, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return RGX_OK;
  if (b.p) CU(cudaFree(b.p));
  b.p = nullptr; b.cap = 0;
  size_t want = std::max(bytes, (size_t)256);
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) { set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e)); b.p = nullptr; return e == cudaErrorMemoryAllocation ? RGX_ENOMEM : RGX_ECUDA; }
  b.cap = want;
  (void)c;
  return RGX_OK;
}

void free_buf(DevBuf& b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; }

// small device block: [0] err flags (int), [8..] totals / flags (8-byte slots)
struct Small {
  int* err;
  unsigned long long* slots;  // 16 slots
};
Small small_of(rgx_ctx* c) { return Small{(int*)c->small.p, (unsigned long long*)((char*)c->small.p + 64)}; }

int get_image(rgx_ctx* c, const rgx_program* p, const DeviceImage** out) {
  std::lock_guard<std::mutex> g(p->mu);
  if ((size_t)c->device >= p->images.size()) p->images.resize(c->device + 1, nullptr);
  DeviceImage* im = p->images[c->device];
  if (!im) {
    im = new DeviceImage();
    std::vector<uint32_t> words;
    pack_program(p->prog, words, im->meta);
    im->in_smem = words.size() * 4 <= SMEM_IMAGE_LIMIT;
    cudaError_t e = cudaMalloc(&im->d_words, words.size() * 4);
    if (e != cudaSuccess) { delete im; set_error(std::string("cudaMalloc(image): ") + cudaGetErrorString(e)); return RGX_ECUDA; }
    e = cudaMemcpy(im->d_words, words.data(), words.size() * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(im->d_words); delete im; set_error(std::string("cudaMemcpy(image): ") + cudaGetErrorString(e)); return RGX_ECUDA; }
    p->images[c->device] = im;
  }
  *out = im;
  return RGX_OK;
}

int check_caps(const DeviceImage* im, bool find) {
  const DevMeta& m = im->meta;
  if (m.num_cap > MAX_CAPS || (find && m.find_engine == FIND_TDFA && m.t_ntags > MAX_CAPS)) {
    set_error("pattern has more capture groups than the device engines support (2*(k+1) <= 32)");
    return RGX_EUNSUPPORTED;
  }
  if (m.n_inst > 65535) { set_error("program too large for the device engines"); return RGX_EUNSUPPORTED; }
  return RGX_OK;
}

template <class K>
int occupancy_grid(rgx_ctx* c, K kernel, int block, size_t smem, int* grid) {
  // static + dynamic shared memory above 48 KiB needs the opt-in, so always declare the dynamic size
  if (smem > 0) CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem));
  if (per_sm < 1) per_sm = 1;
  *grid = per_sm * c->sm_count;
  return RGX_OK;
}

// Size the per-thread backtracking scratch for `threads` machines over inputs of at most max_len
// bytes.  stack_cap == 0 asks for the provable worst case (one entry per (Alt|Capture, offset)).
int plan_scratch(rgx_ctx* c, const DevMeta& m, bool find, uint32_t threads, uint64_t max_len, uint32_t stack_cap,
                 bool visited_for_len, ScratchPlan* sp) {
  std::memset(sp, 0, sizeof(*sp));
  sp->stride = threads;
  const bool needs_bt = (m.flags & F_NEEDS_BT) != 0;
  const bool per_capture = find && (m.flags & F_PER_CAPTURE);
  const bool memo = find ? (m.flags & F_FIND_MEMO) != 0 : (m.flags & F_MATCH_MEMO) != 0;
  const bool bt_engine = find ? m.find_engine == FIND_BT : m.match_engine == MATCH_BT;
  if (!bt_engine) return RGX_OK;
  if (needs_bt || per_capture) {
    uint64_t worst = ((uint64_t)m.n_alt + (per_capture ? (uint64_t)m.n_capinst : 0)) * (max_len + 1) + 4;
    uint64_t cap = stack_cap ? std::min<uint64_t>(stack_cap, worst) : worst;
    if (cap > 0x7FFFFFFFull) { set_error("backtracking stack bound too large"); return RGX_ENOMEM; }
    sp->stack_cap = (uint32_t)cap;
    int rc = ensure(c, c->stack, (size_t)cap * threads * sizeof(uint2));
    if (rc) return rc;
    sp->stack = (uint2*)c->stack.p;
    if (find && !per_capture) {
      sp->cstack_cap = (uint32_t)cap;
      rc = ensure(c, c->cstack, (size_t)cap * m.num_cap * threads * sizeof(int32_t));
      if (rc) return rc;
      sp->cstack = (int32_t*)c->cstack.p;
    }
  }
  if (memo && visited_for_len) {
    uint64_t words = ((uint64_t)m.n_inst * (max_len + 1) + 31) / 32;
    if (words * threads * 4 > (1ull << 36)) { set_error("memoisation bit-vector too large for this input"); return RGX_ENOMEM; }
    sp->visited_words = (uint32_t)words;
    int rc = ensure(c, c->visited, (size_t)words * threads * 4);
    if (rc) return rc;
    sp->visited = (uint32_t*)c->visited.p;
  }
  return RGX_OK;
}

int read_small(rgx_ctx* c, size_t bytes) {  // device small block -> pinned host copy, synchronised
  CU(cudaMemcpyAsync(c->h_small, c->small.p, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return RGX_OK;
}

std::string err_bits(int e) {
  std::string s;
  if (e & ERR_STACK) s += " backtrack-stack";
  if (e & ERR_CSTACK) s += " capture-stack";
  if (e & ERR_VISITED) s += " visited";
  if (e & ERR_RANGE) s += " offset-range";
  if (e & ERR_SLAB) s += " record-slab";
  if (e & ERR_DENSE) s += " dense-input";
  if (e & ERR_HALO) s += " halo";
  if (e & ERR_INTERNAL) s += " internal-invariant";
  return s;
}

}  // namespace

namespace rgx {
void release_device_program(rgx_program* p) {
  for (DeviceImage* im : p->images) {
    if (!im) continue;
    if (im->d_words) cudaFree(im->d_words);
    delete im;
  }
  p->images.clear();
}
}  // namespace rgx

// ---- batched MatchBytes for many programs in ONE launch -------------------------------------------------
template <typename OFFT>
static int match_multi_dev(rgx_ctx* c, const rgx_program* const* progs, uint32_t n_progs, const uint8_t* d_bytes, const OFFT* d_offs,
                           const uint64_t* prog_first, uint8_t* d_out) {
  if (!c || !progs || !prog_first || n_progs == 0) { set_error("null argument"); return RGX_EINVAL; }
  const uint64_t n = prog_first[n_progs];
  for (uint32_t p = 0; p < n_progs; p++)
    if (!progs[p] || prog_first[p] > prog_first[p + 1]) { set_error("rgx_match_multi: bad program table"); return RGX_EINVAL; }
  if (n == prog_first[0]) return RGX_OK;
  CU(cudaSetDevice(c->device));
  // table: metas | image pointers | item_base | prog_first
  std::vector<const DeviceImage*> ims(n_progs);
  bool any_bt = false;
  for (uint32_t p = 0; p < n_progs; p++) {
    int rc = get_image(c, progs[p], &ims[p]);
    if (rc) return rc;
    if ((rc = check_caps(ims[p], false))) return rc;
    any_bt = any_bt || ims[p]->meta.match_engine == MATCH_BT;
  }
  // items: runs of tiles of one program, sized so that a few CTAs per SM cover the batch
  uint64_t total_tiles = 0;
  for (uint32_t p = 0; p < n_progs; p++) total_tiles += (prog_first[p + 1] - prog_first[p] + MM_TILE - 1) / MM_TILE;
  const uint64_t target_items = (uint64_t)c->sm_count * 96;   // fine enough that the slowest program's items do not form a tail
  const uint32_t tiles_per_item = (uint32_t)std::max<uint64_t>(1, (total_tiles + target_items - 1) / target_items);
  std::vector<uint32_t> item_base(n_progs + 1, 0);
  for (uint32_t p = 0; p < n_progs; p++) {
    const uint64_t tiles = (prog_first[p + 1] - prog_first[p] + MM_TILE - 1) / MM_TILE;
    const uint64_t items = (tiles + tiles_per_item - 1) / tiles_per_item;
    if (item_base[p] + items > 0x7FFFFFFFull) { set_error("rgx_match_multi: batch too large"); return RGX_EINVAL; }
    item_base[p + 1] = item_base[p] + (uint32_t)items;
  }
  const size_t off_ptr = ((size_t)n_progs * sizeof(DevMeta) + 7) & ~(size_t)7, off_item = off_ptr + (size_t)n_progs * 8,
               off_first = (off_item + (size_t)(n_progs + 1) * 4 + 7) & ~(size_t)7, tab_bytes = off_first + (size_t)(n_progs + 1) * 8;
  std::vector<uint8_t> tab(tab_bytes, 0);
  for (uint32_t p = 0; p < n_progs; p++) {
    std::memcpy(&tab[(size_t)p * sizeof(DevMeta)], &ims[p]->meta, sizeof(DevMeta));
    const uint32_t* dp = ims[p]->d_words;
    std::memcpy(&tab[off_ptr + (size_t)p * 8], &dp, 8);
  }
  std::memcpy(&tab[off_item], item_base.data(), (size_t)(n_progs + 1) * 4);
  std::memcpy(&tab[off_first], prog_first, (size_t)(n_progs + 1) * 8);
  int rc;
  if (tab != c->mm_key || !c->mm_tab.p) {
    if ((rc = ensure(c, c->mm_tab, tab_bytes))) return rc;
    CU(cudaMemcpyAsync(c->mm_tab.p, tab.data(), tab_bytes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));   // (pageable source: the copy must be over before `tab` goes away)
    c->mm_key = tab;
  }
  MultiArgs a;
  a.metas = (const DevMeta*)c->mm_tab.p;
  a.images = (const uint32_t* const*)((const char*)c->mm_tab.p + off_ptr);
  a.item_base = (const uint32_t*)((const char*)c->mm_tab.p + off_item);
  a.prog_first = (const unsigned long long*)((const char*)c->mm_tab.p + off_first);
  a.n_progs = n_progs; a.n_items = item_base[n_progs]; a.tiles_per_item = tiles_per_item;
  a.image_bytes = 16;
  for (uint32_t p = 0; p < n_progs; p++) {
    const uint32_t b = ims[p]->meta.match_words * 4u;
    if (b <= MM_IMAGE_BYTES) a.image_bytes = std::max(a.image_bytes, (b + 15u) & ~15u);
  }

  Small sm = small_of(c);
  const size_t smem = (size_t)a.image_bytes + MM_TILE_BYTES;
  // first try: stack and visited bits in local memory (no sizing pass, no scratch)
  {
    auto kern = match_multi_kernel<OFFT, true>;
    int grid = 0;
    if ((rc = occupancy_grid(c, kern, MM_TILE, smem, &grid))) return rc;
    if ((uint32_t)grid > a.n_items) grid = (int)a.n_items;
    ScratchPlan sp;
    std::memset(&sp, 0, sizeof sp);
    sp.stride = (uint32_t)grid * MM_TILE;
    CU(cudaMemsetAsync(sm.err, 0, sizeof(int), c->stream));
    kern<<<grid, MM_TILE, smem, c->stream>>>(a, d_bytes, d_offs, d_out, sp, sm.err);
    c->launches++;
    CU(cudaGetLastError());
    if (!any_bt) return RGX_OK;     // nothing can overflow: no readback, the call stays asynchronous
    if ((rc = read_small(c, 64))) return rc;
    const int e = *(int*)c->h_small;
    if (e == 0) return RGX_OK;
    if (e & ~(ERR_STACK | ERR_VISITED)) { set_error("device engine failure:" + err_bits(e)); return RGX_ENOMEM; }
  }
  // some input outgrew the local arrays: the whole batch again with global scratch sized for the longest input
  uint64_t max_len = 0;
  CU(cudaMemsetAsync(c->small.p, 0, 256, c->stream));
  max_len_multi_kernel<OFFT><<<std::max(1, c->sm_count * 4), 256, 0, c->stream>>>(d_offs + prog_first[0], n - prog_first[0], sm.slots);
  c->launches++;
  if ((rc = read_small(c, 256))) return rc;
  max_len = *(unsigned long long*)((char*)c->h_small + 64);
  auto kern = match_multi_kernel<OFFT, false>;
  int grid = 0;
  if ((rc = occupancy_grid(c, kern, MM_TILE, smem, &grid))) return rc;
  if ((uint32_t)grid > a.n_items) grid = (int)a.n_items;
  for (int attempt = 0; attempt < 2; attempt++) {
    // one scratch plan for the whole launch: the largest need over the programs
    ScratchPlan sp;
    std::memset(&sp, 0, sizeof sp);
    sp.stride = (uint32_t)grid * MM_TILE;
    const uint32_t cap = attempt == 0 ? (uint32_t)std::min<uint64_t>(2 * max_len + 64, 1u << 20) : 0;
    for (uint32_t p = 0; p < n_progs; p++) {
      ScratchPlan one;
      rc = plan_scratch(c, ims[p]->meta, false, sp.stride, max_len, cap, true, &one);
      if (rc) break;
      sp.stack_cap = std::max(sp.stack_cap, one.stack_cap); sp.visited_words = std::max(sp.visited_words, one.visited_words);
    }
    if (rc == RGX_ENOMEM && grid > c->sm_count) { grid = c->sm_count; attempt--; continue; }
    if (rc) return rc;
    // (plan_scratch grows the buffers monotonically, so after the loop they hold the largest request)
    sp.stack = (uint2*)c->stack.p; sp.visited = (uint32_t*)c->visited.p;
    CU(cudaMemsetAsync(sm.err, 0, sizeof(int), c->stream));
    kern<<<grid, MM_TILE, smem, c->stream>>>(a, d_bytes, d_offs, d_out, sp, sm.err);
    c->launches++;
    CU(cudaGetLastError());
    if ((rc = read_small(c, 64))) return rc;
    const int e = *(int*)c->h_small;
    if (e == 0) return RGX_OK;
    if (attempt == 1 || !(e & (ERR_STACK | ERR_CSTACK))) { set_error("device engine scratch exhausted:" + err_bits(e)); return RGX_ENOMEM; }
  }
  return RGX_OK;
}

extern "C" {

void rgx_ctx_destroy(rgx_ctx* c);

int rgx_ctx_create(int32_t device, rgx_ctx** out) {
  if (!out) { set_error("rgx_ctx_create: null argument"); return RGX_EINVAL; }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error(std::string("no CUDA device available (this library has no CPU matching path): ") + cudaGetErrorString(e));
    return RGX_ECUDA;
  }
  if (device < 0 || device >= n) { set_error("rgx_ctx_create: bad device ordinal"); return RGX_EINVAL; }
  CU(cudaSetDevice(device));
  auto* c = new rgx_ctx();
  c->device = device;
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  if (prop.major < 10) {
    delete c;
    set_error("regengo_b200 is built for sm_100a only");
    return RGX_ECUDA;
  }
  // (a half-built context is torn down by rgx_ctx_destroy: it frees whatever exists so far)
  auto fail = [&](cudaError_t e, const char* what) { set_error(std::string(what) + ": " + cudaGetErrorString(e)); rgx_ctx_destroy(c); return RGX_ECUDA; };
  cudaError_t e2;
  if ((e2 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e2, "cudaStreamCreate");
  if ((e2 = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e2, "cudaStreamCreate");
  if ((e2 = cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e2, "cudaStreamCreate");
  if ((e2 = cudaMallocHost(&c->h_small, 4096)) != cudaSuccess) return fail(e2, "cudaMallocHost");
  int rc = ensure(c, c->small, 4096);
  if (rc) { rgx_ctx_destroy(c); return rc; }
  if ((e2 = cudaMemsetAsync(c->small.p, 0, 4096, c->stream)) != cudaSuccess) return fail(e2, "cudaMemsetAsync");
  *out = c;
  return RGX_OK;
}

void rgx_ctx_destroy(rgx_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  DevBuf* bufs[] = {&c->stack, &c->cstack, &c->visited, &c->small, &c->in_bytes, &c->in_offs, &c->out_flag, &c->out_rec,
                    &c->out_reps, &c->out_aux, &c->mm_tab, &c->rp_tmpl, &c->rp_len, &c->rp_offs, &c->rp_out, &c->fa_count, &c->fa_keys, &c->fa_caps, &c->fa_reps, &c->ch_a, &c->ch_b,
                    &c->ch_sel, &c->ch_reps, &c->ch_selbase, &c->ch_repsbase, &c->ch_segsel, &c->ch_segreps, &c->ch_entry, &c->ch_tile};
  for (DevBuf* b : bufs) free_buf(*b);
  if (c->h_small) cudaFreeHost(c->h_small);
  for (auto& e : c->ev) if (e) cudaEventDestroy(e);
  for (auto& e : c->chunk_ev) if (e) cudaEventDestroy(e);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int64_t rgx_ctx_launches(const rgx_ctx* c) { return c ? c->launches : 0; }
int64_t rgx_ctx_stat(const rgx_ctx* c, int32_t which) { return c && which == 0 ? c->stat_seq_chunks : -1; }

int rgx_ctx_set_chunk_bytes(rgx_ctx* c, uint64_t bytes) {
  if (!c || bytes < (1u << 16) || (bytes & 0x7FFFu)) { set_error("rgx_ctx_set_chunk_bytes: need a multiple of 32 KiB, at least 64 KiB"); return RGX_EINVAL; }
  c->chunk_bytes = bytes;
  return RGX_OK;
}

int rgx_ctx_enable_timing(rgx_ctx* c, int32_t on) {
  if (!c) return RGX_EINVAL;
  CU(cudaSetDevice(c->device));
  if (on && !c->ev[0]) for (auto& e : c->ev) CU(cudaEventCreate(&e));
  c->timing = on != 0;
  return RGX_OK;
}
int rgx_ctx_last_timing(const rgx_ctx* c, float* out_ms) {
  if (!c || !out_ms) return RGX_EINVAL;
  for (int i = 0; i < 4; i++) out_ms[i] = c->last_ms[i];
  return RGX_OK;
}
void* rgx_ctx_stream(const rgx_ctx* c) { return c ? (void*)c->stream : nullptr; }
int rgx_ctx_sync(rgx_ctx* c) {
  if (!c) return RGX_EINVAL;
  CU(cudaStreamSynchronize(c->stream));
  return RGX_OK;
}

int rgx_dev_alloc(rgx_ctx* c, size_t bytes, void** out) {
  if (!c || !out) return RGX_EINVAL;
  CU(cudaSetDevice(c->device));
  cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
  if (e != cudaSuccess) { set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e)); return RGX_ENOMEM; }
  return RGX_OK;
}
int rgx_dev_free(rgx_ctx* c, void* p) {
  if (!c) return RGX_EINVAL;
  CU(cudaSetDevice(c->device));
  CU(cudaFree(p));
  return RGX_OK;
}
int rgx_dev_upload(rgx_ctx* c, void* d, const void* h, size_t bytes) {
  if (!c) return RGX_EINVAL;
  CU(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return RGX_OK;
}
int rgx_dev_download(rgx_ctx* c, void* h, const void* d, size_t bytes) {
  if (!c) return RGX_EINVAL;
  CU(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return RGX_OK;
}

// ---- batched MatchBytes / FindBytes -------------------------------------------------------------------
static int batch_dev(rgx_ctx* c, const rgx_program* p, int what, const uint8_t* d_bytes, const uint64_t* d_offs, uint64_t n,
                     uint8_t* d_flag, int64_t* d_rec) {
  if (!c || !p) { set_error("null context/program"); return RGX_EINVAL; }
  if (n == 0) return RGX_OK;
  CU(cudaSetDevice(c->device));
  const DeviceImage* im;
  int rc = get_image(c, p, &im);
  if (rc) return rc;
  const DevMeta& m = im->meta;
  if (what == 1 && m.find_engine == FIND_NONE) {
    set_error("FindBytes is not generated for a pattern without capture groups (regengo.go:110)");
    return RGX_EUNSUPPORTED;
  }
  rc = check_caps(im, what == 1);
  if (rc) return rc;
  Small sm = small_of(c);
  // longest input (sizes the per-thread scratch)
  CU(cudaMemsetAsync(c->small.p, 0, 256, c->stream));
  max_len_kernel<<<std::max(1, c->sm_count), 256, 0, c->stream>>>(d_offs, n, sm.slots);
  c->launches++;
  rc = read_small(c, 256);
  if (rc) return rc;
  const uint64_t max_len = *(unsigned long long*)((char*)c->h_small + 64);
  const size_t smem = im->in_smem ? (size_t)m.image_words * 4 : 0;
  int grid = 0;
  if (what == 0) rc = occupancy_grid(c, batch_kernel<0>, 256, smem, &grid);
  else rc = occupancy_grid(c, batch_kernel<1>, 256, smem, &grid);
  if (rc) return rc;
  const uint64_t need_blocks = (n + 255) / 256;
  if ((uint64_t)grid > need_blocks) grid = (int)need_blocks;
  for (int attempt = 0; attempt < 2; attempt++) {
    ScratchPlan sp;
    const uint32_t cap = attempt == 0 ? (uint32_t)std::min<uint64_t>(2 * max_len + 64, 1u << 20) : 0;
    rc = plan_scratch(c, m, what == 1, (uint32_t)grid * 256u, max_len, cap, true, &sp);
    if (rc == RGX_ENOMEM && grid > c->sm_count) { grid = c->sm_count; attempt--; continue; }
    if (rc) return rc;
    CU(cudaMemsetAsync(sm.err, 0, sizeof(int), c->stream));
    if (what == 0)
      batch_kernel<0><<<grid, 256, smem, c->stream>>>(m, im->d_words, im->in_smem ? 1 : 0, d_bytes, d_offs, n, d_flag, d_rec, sp, sm.err);
    else
      batch_kernel<1><<<grid, 256, smem, c->stream>>>(m, im->d_words, im->in_smem ? 1 : 0, d_bytes, d_offs, n, d_flag, d_rec, sp, sm.err);
    c->launches++;
    CU(cudaGetLastError());
    rc = read_small(c, 64);
    if (rc) return rc;
    const int e = *(int*)c->h_small;
    if (e == 0) return RGX_OK;
    if (attempt == 1 || !(e & (ERR_STACK | ERR_CSTACK))) {
      set_error("device engine scratch exhausted:" + err_bits(e));
      return RGX_ENOMEM;
    }
  }
  return RGX_OK;
}

int rgx_match_batch_dev(rgx_ctx* c, const rgx_program* p, const uint8_t* d_bytes, const uint64_t* d_offs, uint64_t n, uint8_t* d_out) {
  if (!p) { set_error("null context/program"); return RGX_EINVAL; }
  if (n == 0) return RGX_OK;
  const uint64_t pf[2] = {0, n};
  return match_multi_dev<unsigned long long>(c, &p, 1, d_bytes, (const unsigned long long*)d_offs, pf, d_out);
}
int rgx_find_batch_dev(rgx_ctx* c, const rgx_program* p, const uint8_t* d_bytes, const uint64_t* d_offs, uint64_t n,
                       uint8_t* d_found, int64_t* d_out) {
  return batch_dev(c, p, 1, d_bytes, d_offs, n, d_found, d_out);
}

static int batch_host(rgx_ctx* c, const rgx_program* p, int what, const uint8_t* bytes, const uint64_t* offs, uint64_t n,
                      uint8_t* flag, int64_t* rec) {
  if (!c || !p || (n && (!offs || !flag))) { set_error("null argument"); return RGX_EINVAL; }
  if (n == 0) return RGX_OK;
  if (what == 1 && p->prog.find_engine == FIND_NONE) {
    set_error("FindBytes is not generated for a pattern without capture groups (regengo.go:110)");
    return RGX_EUNSUPPORTED;
  }
  CU(cudaSetDevice(c->device));
  const uint64_t base = offs[0], total = offs[n] - base;
  int rc;
  if ((rc = ensure(c, c->in_bytes, total + 16))) return rc;
  if ((rc = ensure(c, c->in_offs, (n + 1) * 8))) return rc;
  if ((rc = ensure(c, c->out_flag, n))) return rc;
  const int nc = p->prog.prog.num_cap;
  if (what == 1 && (rc = ensure(c, c->out_rec, n * nc * 8))) return rc;
  if (total) CU(cudaMemcpyAsync(c->in_bytes.p, bytes + base, total, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->in_offs.p, offs, (n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  // offsets are used as given; shift the byte pointer so that offs[0] maps to the copied base
  const uint8_t* d_bytes = (const uint8_t*)c->in_bytes.p - base;
  rc = batch_dev(c, p, what, d_bytes, (const uint64_t*)c->in_offs.p, n, (uint8_t*)c->out_flag.p, (int64_t*)c->out_rec.p);
  if (rc) return rc;
  CU(cudaMemcpyAsync(flag, c->out_flag.p, n, cudaMemcpyDeviceToHost, c->stream));
  if (what == 1) CU(cudaMemcpyAsync(rec, c->out_rec.p, n * nc * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return RGX_OK;
}

int rgx_match_batch(rgx_ctx* c, const rgx_program* p, const uint8_t* bytes, const uint64_t* offs, uint64_t n, uint8_t* out) {
  return batch_host(c, p, 0, bytes, offs, n, out, nullptr);
}
int rgx_find_batch(rgx_ctx* c, const rgx_program* p, const uint8_t* bytes, const uint64_t* offs, uint64_t n, uint8_t* found,
                   int64_t* out) {
  if (n && !out) { set_error("null argument"); return RGX_EINVAL; }
  return batch_host(c, p, 1, bytes, offs, n, found, out);
}

int rgx_match_multi_dev(rgx_ctx* c, const rgx_program* const* progs, uint32_t n_progs, const uint8_t* d_bytes, const uint32_t* d_offs32,
                        const uint64_t* prog_first, uint8_t* d_out) {
  return match_multi_dev<uint32_t>(c, progs, n_progs, d_bytes, d_offs32, prog_first, d_out);
}

int rgx_match_multi(rgx_ctx* c, const rgx_program* const* progs, uint32_t n_progs, const uint8_t* bytes, const uint64_t* offs,
                    const uint64_t* prog_first, uint8_t* out) {
  if (!c || !progs || !prog_first || !offs || n_progs == 0) { set_error("null argument"); return RGX_EINVAL; }
  const uint64_t i0 = prog_first[0], n = prog_first[n_progs] - i0;
  if (n == 0) return RGX_OK;
  if (!out) { set_error("null argument"); return RGX_EINVAL; }
  const uint64_t base = offs[i0], total = offs[i0 + n] - base;
  CU(cudaSetDevice(c->device));
  int rc;
  if ((rc = ensure(c, c->in_bytes, total + 16))) return rc;
  if ((rc = ensure(c, c->in_offs, (n + 1) * 8))) return rc;
  if ((rc = ensure(c, c->out_flag, n))) return rc;
  // The ABI's 64-bit offsets are uploaded as they are (no host pass over them) and the kernel's 64-bit-offset form
  // reads them directly; the byte pointer is shifted so that offs[i0] maps to the first uploaded byte.
  std::vector<uint64_t> pf(n_progs + 1);
  for (uint32_t p = 0; p <= n_progs; p++) pf[p] = prog_first[p] - i0;
  if (total) CU(cudaMemcpyAsync(c->in_bytes.p, bytes + base, total, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->in_offs.p, offs + i0, (n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  rc = match_multi_dev<unsigned long long>(c, progs, n_progs, (const uint8_t*)c->in_bytes.p - base, (const unsigned long long*)c->in_offs.p, pf.data(),
                                           (uint8_t*)c->out_flag.p);
  if (rc) return rc;
  CU(cudaMemcpyAsync(out, c->out_flag.p, n, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return RGX_OK;
}

}  // extern "C"

#include "capi_findall.inc"
#include "capi_stream.inc"
#include "capi_replace.inc"

extern "C" {

int64_t rgx_program_device_plan(const rgx_program* p, char* buf, size_t cap) {
  if (!p) { set_error("null argument"); return RGX_EINVAL; }
  std::vector<uint32_t> words;
  DevMeta m;
  pack_program(p->prog, words, m);
  std::string s = "{";
  auto kv = [&](const char* k, long long v) { if (s.size() > 1) s += ", "; s += "\""; s += k; s += "\": "; s += std::to_string(v); };
  kv("image_bytes", (long long)m.image_words * 4);
  kv("match_image_bytes", (long long)m.match_words * 4);
  kv("find_engine", m.find_engine);
  kv("gen_kind", m.gen_kind);
  kv("prefix_len", m.prefix_len);
  kv("nullable", m.nullable);
  kv("fast_tdfa_scan", fast_tdfa_scan_ok(m) && parallel_findall_ok(m, p->prog) ? 1 : 0);
  kv("parallel_findall", parallel_findall_ok(m, p->prog) ? 1 : 0);
  kv("scan6_image_bytes", (long long)m.w6_words * 4);
  kv("scan6_descriptors", m.w6_ndesc);
  kv("w6_off", m.w6_off); kv("w6_desc", m.w6_desc); kv("w6_fent", m.w6_fent); kv("w6_init", m.w6_init); kv("w6_maxev", m.w6_maxev);
  kv("tdfa_states", m.t_ns); kv("tdfa_tags", m.t_ntags); kv("tdfa_start_any", m.t_start_any); kv("tdfa_n_init_any", m.t_n_init_any);
  kv("run_anchor", m.run_ok);
  kv("run_literal", m.run_ok ? m.run_lit : -1);
  kv("run_linear_elements", m.lin_n);
  kv("straight_line_steps", m.sl_n);
  kv("straight_line_prefix_steps", m.slp_n);
  kv("linear_prefix_findall_scan", (!fast_tdfa_scan_ok(m) && !(m.find_engine == FIND_BT && m.run_ok) && m.find_engine == FIND_BT && m.sl_n == 0 && m.slp_n >= 2 && m.slp_n <= 32) ? 1 : 0);
  kv("linear_findall_scan", (!fast_tdfa_scan_ok(m) && !(m.find_engine == FIND_BT && m.run_ok) && m.find_engine == FIND_BT && m.sl_n > 0 && m.sl_n <= 32 && m.sl_caps_ok) ? 1 : 0);
  kv("straight_line_classes", m.sl_ncls);
  kv("sl_cm_off", m.off_sl_cm);
  kv("sl_caps_ok", m.sl_caps_ok);
  {
    // class index of every step, capture offsets (hex, one byte each): what the bit-plane kernels read
    s += ", \"sl_cls\": \"";
    for (int i = 0; i < m.slp_n && i < 32; i++) { char h[8]; std::snprintf(h, sizeof h, "%02x", m.sl_cls[i]); s += h; }
    s += "\", \"sl_cap\": \"";
    for (int i = 0; i < m.num_cap && i < 32 && m.sl_caps_ok; i++) { char h[8]; std::snprintf(h, sizeof h, "%02x", m.sl_cap[i]); s += h; }
    s += "\"";
  }
  kv("n_alt", m.n_alt);
  kv("n_empty", m.n_empty);
  s += ", \"prefix\": \"";
  for (int i = 0; i < m.prefix_len; i++) { char h[8]; std::snprintf(h, sizeof h, "%02x", m.prefix_bytes[i]); s += h; }
  s += "\"}";
  if (buf && cap > s.size()) std::memcpy(buf, s.c_str(), s.size() + 1);
  return (int64_t)s.size();
}

// The packed device image itself (u32 words), for host-side tests of the packing: returns the number of words.
int64_t rgx_program_device_image(const rgx_program* p, uint32_t* words_out, size_t cap_words) {
  if (!p) { set_error("null argument"); return RGX_EINVAL; }
  std::vector<uint32_t> words;
  DevMeta m;
  pack_program(p->prog, words, m);
  if (words_out && cap_words >= words.size()) std::memcpy(words_out, words.data(), words.size() * 4);
  return (int64_t)words.size();
}

}  // extern "C"
