// Device-side form of one compiled pattern: a packed table image (u32 words, 16-byte aligned)
// that every CTA stages from HBM/L2 into shared memory with ONE TMA bulk copy
// (cp.async.bulk.shared::cluster.global + mbarrier complete_tx), plus a small by-value DevMeta
// with the scalars and the section offsets.
//
// The image holds what the reference bakes into generated Go source: the instruction list of the
// goto-machine (internal/compiler/instructions.go), the class bitmaps (charclass.go:43-74), the
// Thompson closure masks (thompson.go:133-157) and the TDFA tables (tdfa.go:547-794), re-packed
// for the GPU: 16-byte instruction records, u32 transition cells carrying the next state and the
// index of a de-duplicated tag-action list.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "blob.hpp"

namespace rgx {

constexpr int MAX_CAPS = 32;         // 2*(k+1) <= 32 on the device paths
constexpr int MAX_PREFIX = 8;
constexpr uint32_t TDFA_NONE = 0xFFFFu;
constexpr uint32_t S6_DEAD = 0xFFC00000u;    // scan6 cell without a transition (every cell >= S6_DEAD is dead)
constexpr uint32_t S6_EVMIN = 0x00400000u;   // scan6 cells >= S6_EVMIN (and < S6_DEAD) are events
constexpr uint32_t S6_ACC = 1u << 20, S6_ACC_EOT = 1u << 21;   // scan6 descriptor flags
constexpr uint32_t S6_FENT = 4;                  // tag entries per scan6 descriptor
constexpr uint32_t S6_ENT_ACCEPT = 0x80000000u;  // scan6 tag entry: an accept action of the event's next state
constexpr uint32_t S6_IMAGE_LIMIT = 96 * 1024;   // bytes of walk image staged per CTA
constexpr uint32_t SMEM_IMAGE_LIMIT = 160 * 1024;

enum GenKind : int { GEN_ALL = 0, GEN_BYTESET = 1, GEN_PREFIX = 2 };
enum LinKind : uint32_t { LIN_CAP = 0, LIN_LIT = 1, LIN_CLS = 2, LIN_LOOP = 3, LIN_MATCH = 4 };

struct DevMeta {
  int32_t n_inst, start, num_cap, flags, prefix, match_engine, find_engine;
  uint32_t image_words;        // multiple of 4
  uint32_t match_words;        // multiple of 4: the leading part of the image that the MatchBytes engines read
  uint32_t off_inst, off_cls, off_th_eps, off_th_cond, off_rng_idx, off_rng_pairs;
  uint32_t off_t_trans, off_t_accept, off_t_alist_off, off_t_alist, off_t_init, off_first;
  // scan6 walk image (kernels_scan6.cuh; w6_ok = 0: absent): the contiguous sub-range [w6_off, w6_off + w6_words) of the
  // image is all the FindAll scan stages into shared memory.  Offsets below are in words, relative to w6_off:
  //   rows   t_ns x 256 cells at 0: idx:10 << 22 | byte offset of the next state's row.  idx 0 = CHEAP step: no
  //          transition tag list fires and the step is either a self-loop or a move between two states that accept
  //          neither in the text nor at its end.  idx 0x3FF (S6_DEAD) = no transition (bytes >= 128 included,
  //          tdfa.go:944-946).  Any other idx = EVENT, described by desc[idx].
  //   desc   8 words per descriptor (w6_fent = w6_desc + 1):
  //          [0] accepts:1 (S6_ACC) | accepts at EOT:1 (S6_ACC_EOT) | n:3 << 24 -- flags of the event's next state and the
  //              number of tag entries
  //          [1..4] the event's tag actions in application order: tag * 128 | offset:8 << 16, S6_ENT_ACCEPT set on the next
  //              state's accept actions (applied at the end of the state's run, and only while that state accepts); the
  //              others are the transition's actions (applied at the step's position)
  //   init   tags set to the start position by initialTagsAny
  uint32_t w6_off, w6_words, w6_desc, w6_fent, w6_init;
  int32_t w6_ok, w6_ndesc;
  // (start filter of scan6: up to three of the first four bytes of the literal prefix -- necessary, not sufficient;
  // the walk starts in startStateAny at the candidate start)  Without a literal first byte: the first-byte set as
  // w6_nrng <= 2 ASCII ranges [w6_rlo, w6_rhi].
  int32_t w6_nrng;
  int32_t w6_maxev;   // most events one walk can log (255: unbounded): sizes the per-candidate log of scan6
  uint8_t w6_rlo[2], w6_rhi[2];
  int32_t t_ns, t_ntags, t_start_begin, t_start_any, t_n_init_begin, t_n_init_any;
  uint32_t th_start_lo, th_start_hi, th_accept_lo, th_accept_hi, th_char_lo, th_char_hi;
  // FindAll candidate generator (start filter); see findall_kernels.cu
  int32_t gen_kind, prefix_len;
  uint8_t prefix_bytes[MAX_PREFIX];
  int32_t nullable;            // a match attempt can succeed without consuming input
  int32_t n_alt;               // number of Alt instructions (stack sizing)
  int32_t n_capinst;
  int32_t n_empty;             // number of EmptyWidth instructions (their outcome depends on the slice origin)
  // "leading class loop + literal" start filter of the backtracking FindAll scan (kernels_btrun.cuh):
  // the program is  (cap|nop)* C+ (cap|nop)* b ...  with b a byte outside the class C
  int32_t run_ok;              // shape recognised
  int32_t run_class_pc;        // the Rune instruction of C (its 256-bit set is at off_cls + 8*pc)
  int32_t run_lit;             // b
  uint32_t run_start_caps;     // capture slots written before the loop (they hold the attempt's start)
  int32_t run_resume_pc;       // instruction after the loop (the Alt's exit): attempts of a run resume here at offset p
  // When everything after the leading loop is a straight line of captures, literal bytes, ASCII classes and
  // ATOMIC greedy class loops (device_program.cu), the attempt has exactly one path that can succeed -- the
  // greedy one -- and is evaluated directly instead of by the goto-machine: lin[i] = kind | arg << 8.
  int32_t lin_n;               // 0: not linear
  uint32_t lin[40];
  // Straight-line WHOLE program (no Alt, no EmptyWidth): step i consumes one byte of class sl_cls[i]; the
  // classes are de-duplicated (<= 8) and cm[c] (256 bytes at off_sl_cm) has bit k set iff byte c is in class k.
  // Lets the FindReader attempt table be computed bit-parallel (kernels_stream.cuh).
  int32_t sl_n;                // 0: not straight-line
  int32_t sl_ncls;
  uint32_t off_sl_cm;
  uint8_t sl_cls[32];
  uint8_t sl_cap[32];          // capture slot -> offset from the match start (valid when sl_caps_ok)
  int32_t sl_caps_ok;
  int32_t slp_n;               // straight-line PREFIX: steps every match passes first (== sl_n for a whole line; < 2: none)
};

struct DeviceImage {
  DevMeta meta;
  uint32_t* d_words = nullptr;  // device copy of the image
  bool in_smem = true;          // image fits the shared-memory budget
};

// host: build image words + meta from a Program (device_program.cu)
void pack_program(const Program& P, std::vector<uint32_t>& words, DevMeta& meta);

#ifdef __CUDACC__
// View over the staged image.
struct DProg {
  const uint32_t* img;
  __device__ __forceinline__ uint4 inst(const DevMeta& m, int pc) const {
    return reinterpret_cast<const uint4*>(img + m.off_inst)[pc];
  }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Stage the image into shared memory with a 1-D TMA bulk copy.  `mbar` is an 8-byte aligned
// shared u64.  All threads of the CTA must call this; returns after the bytes have landed.
__device__ __forceinline__ void stage_image_tma(uint32_t* smem_dst, const uint32_t* gsrc, uint32_t words,
                                                unsigned long long* mbar) {
  const uint32_t bytes = words * 4u;
  const uint32_t bar = smem_u32(mbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar)
        : "memory");
  }
}
#endif

}  // namespace rgx
