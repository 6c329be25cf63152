// ReplaceAllBytesAppend over a BATCH of inputs (internal/compiler/replace.go:192-273).
//
// Per input the generated method is a loop of FindBytesReuse on input[offset:] -- so the skip-restart rule (Q1) and the
// re-anchoring of every slice (Q3) apply -- that locates the match text with bytes.Index from the slice start (Q16),
// appends input[lastEnd:matchStart] and the template expansion, and continues behind the match (one byte further after
// an empty match).  One thread runs that loop for one input, exactly in that order; the expansion reads the capture
// offsets of the TRUE match (relative to the slice) even when bytes.Index located the text earlier.
// Two passes: MODE 0 computes the output length per input, an exclusive scan places the outputs, MODE 1 writes them.
//
// Template image (words): [n_seg, byte offset of the literal pool from the image start, {group or 0xFFFFFFFF, literal
// offset, literal length} x n_seg], then the literal bytes (csrc/replace_template.cpp resolves names and drops
// references to groups the pattern does not have).
#pragma once
#include "engines.cuh"
#include "kernels_stream.cuh"

namespace rgx {

template <int MODE>
__global__ void __launch_bounds__(128) replace_batch_kernel(const DevMeta m, const uint32_t* __restrict__ gimg, const int in_smem,
                                                            const uint8_t* __restrict__ bytes, const uint64_t* __restrict__ offs,
                                                            const uint64_t n, const uint32_t* __restrict__ tmpl,
                                                            unsigned long long* __restrict__ out_len,
                                                            const unsigned long long* __restrict__ out_offs, uint8_t* __restrict__ out_bytes,
                                                            const ScratchPlan sp, int* err) {
  extern __shared__ __align__(16) uint32_t smem_img[];
  __shared__ __align__(8) unsigned long long mbar;
  const uint32_t* img = gimg;
  if (in_smem) { stage_image_tma(smem_img, gimg, m.image_words, &mbar); img = smem_img; }
  const Scratch sc = scratch_of(sp);
  const int nc = m.find_engine == FIND_TDFA ? m.t_ntags : m.num_cap;
  const uint32_t n_seg = tmpl[0];
  const uint8_t* lits = reinterpret_cast<const uint8_t*>(tmpl) + tmpl[1];

  for (uint64_t i = sc.tid; i < n; i += sp.stride) {
    const uint8_t* in = bytes + offs[i];
    const int64_t l = (int64_t)(offs[i + 1] - offs[i]);
    uint8_t* o = MODE == 1 ? out_bytes + out_offs[i] : nullptr;
    unsigned long long w = 0;
    auto put = [&](const uint8_t* src, const int64_t cnt) {
      if (MODE == 1) for (int64_t k = 0; k < cnt; k++) o[w + k] = src[k];
      w += (unsigned long long)cnt;
    };
    int64_t offset = 0, last_end = 0;
    int64_t rec[MAX_CAPS];
    for (;;) {
      const uint8_t* rem = in + offset;
      const int64_t rl = l - offset;
      int found;
      if (m.find_engine == FIND_TDFA) {
        found = tdfa_find(m, img, rem, rl, 0, rec);
      } else {
        int32_t caps[MAX_CAPS];
        found = bt_machine<MODE_FIND>(m, img, rem, rl, 0, caps, sc, err);
        if (found) bt_emit_record(caps, nc, 0, rl, 0, rec);
      }
      if (!found) break;
      const int64_t mlen = rec[1] - rec[0];
      const int64_t midx = index_of_text(rem, 0, rec[0], mlen);   // bytes.Index(remaining, match.Match)
      const int64_t match_start = offset + midx, match_end = match_start + mlen;
      put(in + last_end, match_start - last_end);
      for (uint32_t s = 0; s < n_seg; s++) {
        const uint32_t g = tmpl[2 + 3 * s];
        if (g == 0xFFFFFFFFu) put(lits + tmpl[3 + 3 * s], (int64_t)tmpl[4 + 3 * s]);
        else if (rec[2 * g] >= 0) put(rem + rec[2 * g], rec[2 * g + 1] - rec[2 * g]);
      }
      last_end = match_end;
      if (mlen > 0) offset = match_end;
      else if (match_end < l) offset = match_end + 1;
      else break;
    }
    put(in + last_end, l - last_end);
    if (MODE == 0) out_len[i] = w;
  }
}

}  // namespace rgx
