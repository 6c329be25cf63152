// Device interpreters for the three engines regengo generates per pattern.  Each one executes
// the packed program image (device_program.cuh) with the exact step order of the generated Go:
//   bt_machine<MODE>  the backtracking goto-machine: MatchBytes (compiler.go:740-871,
//                     backtracking.go:9-77), FindBytesReuse (find.go:469-591, backtracking.go:83-165)
//                     and one FindAllBytesAppend attempt (find.go:130-316); instruction bodies
//                     instructions.go:51-605, captures.go:123-158
//   thompson_match    the uint64 state-set loop (thompson.go:69-197)
//   tdfa_walk         one start position of findBytesInternal (tdfa.go:831-994)
// One thread runs one machine sequentially, so restart order, failure offsets (the skip-restart
// rule) and checkpoint/restore order are reproduced by construction.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "device_program.cuh"

namespace rgx {

enum { OP_ALT = 0, OP_ALTMATCH, OP_CAPTURE, OP_EMPTY, OP_MATCH, OP_FAIL, OP_NOP, OP_RUNE, OP_RUNE1, OP_ANY, OP_ANYNOTNL };
enum { EMPTY_BEGIN_LINE = 1, EMPTY_END_LINE = 2, EMPTY_BEGIN_TEXT = 4, EMPTY_END_TEXT = 8, EMPTY_WORD = 16, EMPTY_NOWORD = 32 };
enum { MODE_MATCH = 0, MODE_FIND = 1, MODE_FINDALL = 2 };
enum { ERR_STACK = 1, ERR_CSTACK = 2, ERR_VISITED = 4, ERR_RANGE = 8, ERR_SLAB = 16, ERR_DENSE = 32, ERR_HALO = 64, ERR_INTERNAL = 128 /* an invariant of a parallel formulation did not hold */ };

constexpr int32_t CAP_ZERO = INT32_MIN;  // a capture still holding Go's zero value (absolute 0)

// Per-thread scratch in global memory, interleaved across threads so that lanes at equal depth
// touch neighbouring words.
struct Scratch {
  uint2* stack;        // entry j of thread t: stack[j * stride + t]
  int32_t* cstack;     // checkpoint c, capture k of thread t: cstack[(c * num_cap + k) * stride + t]
  uint32_t* visited;   // word j of thread t: visited[j * stride + t]
  uint32_t stack_cap, cstack_cap, visited_words;
  uint32_t stride, tid;
};

__device__ __forceinline__ bool is_word_byte(uint8_t b) {  // compiler.go:674-688
  return (b >= 'a' && b <= 'z') || (b >= 'A' && b <= 'Z') || (b >= '0' && b <= '9') || b == '_';
}

// unicode/utf8.DecodeRune [go-stdlib]
__device__ __forceinline__ int decode_rune_dev(const uint8_t* p, int64_t n, int32_t& r) {
  r = 0xFFFD;
  if (n < 1) return 0;
  uint32_t b0 = p[0];
  if (b0 < 0x80) { r = (int32_t)b0; return 1; }
  if (b0 < 0xC2) return 1;
  if (b0 < 0xE0) {
    if (n < 2 || (p[1] & 0xC0) != 0x80) return 1;
    r = (int32_t)(((b0 & 0x1F) << 6) | (p[1] & 0x3F));
    return 2;
  }
  if (b0 < 0xF0) {
    if (n < 3) return 1;
    uint32_t lo = 0x80, hi = 0xBF;
    if (b0 == 0xE0) lo = 0xA0; else if (b0 == 0xED) hi = 0x9F;
    if (p[1] < lo || p[1] > hi || (p[2] & 0xC0) != 0x80) return 1;
    r = (int32_t)(((b0 & 0x0F) << 12) | ((p[1] & 0x3Fu) << 6) | (p[2] & 0x3F));
    return 3;
  }
  if (b0 < 0xF5) {
    if (n < 4) return 1;
    uint32_t lo = 0x80, hi = 0xBF;
    if (b0 == 0xF0) lo = 0x90; else if (b0 == 0xF4) hi = 0x8F;
    if (p[1] < lo || p[1] > hi || (p[2] & 0xC0) != 0x80 || (p[3] & 0xC0) != 0x80) return 1;
    r = (int32_t)(((b0 & 0x07) << 18) | ((p[1] & 0x3Fu) << 12) | ((p[2] & 0x3Fu) << 6) | (p[3] & 0x3F));
    return 4;
  }
  return 1;
}

__device__ __forceinline__ int64_t index_byte_dev(const uint8_t* p, int64_t from, int64_t l, uint8_t b) {
  for (int64_t i = from; i < l; i++) if (p[i] == b) return i;
  return -1;
}

__device__ __forceinline__ void clear_visited(const Scratch& sc) {
  for (uint32_t j = 0; j < sc.visited_words; j++) sc.visited[(size_t)j * sc.stride + sc.tid] = 0;
}

// The goto-machine.  `in`/`l` are the slice the generated method sees; `base` is the origin of the
// int32 relative encoding of offsets kept on the stack and in caps[] (0 for MATCH/FIND, the attempt's
// searchStart for FINDALL).  caps[] entries are base-relative, CAP_ZERO = Go zero value.
// Returns 1 on Match.  MODE_FINDALL: one attempt at searchStart (0 => "searchStart++").
// TRACK (MODE_FINDALL only): *track_fail = offset of the final failure (the generated FindBytesReuse restarts
// at that offset + 1) and *track_reach = one past the highest input position the attempt looked
// at (byte reads and end-of-input tests): an attempt whose reach is <= l' behaves identically on in[0:l'].
template <int MODE, bool TRACK = false>
__device__ int bt_machine(const DevMeta& m, const uint32_t* __restrict__ img, const uint8_t* __restrict__ in, int64_t l,
                          int64_t search_start, int32_t* caps, const Scratch& sc, int* err, int resume_pc = -1,
                          int64_t resume_off = 0, uint32_t resume_start_caps = 0, int64_t* track_fail = nullptr,
                          int64_t* track_reach = nullptr) {
  const int ncap = m.num_cap;
  const bool anchored = (m.flags & F_ANCHORED) != 0;
  const bool needs_bt = (m.flags & F_NEEDS_BT) != 0;
  const bool per_capture = (m.flags & F_PER_CAPTURE) != 0;
  const bool memo = MODE == MODE_MATCH ? (m.flags & F_MATCH_MEMO) != 0 : (m.flags & F_FIND_MEMO) != 0;
  const bool has_prefix = MODE == MODE_MATCH && (m.flags & F_HAS_PREFIX) && !anchored;
  const int64_t base = MODE == MODE_FINDALL ? search_start : 0;
  const uint4* insts = reinterpret_cast<const uint4*>(img + m.off_inst);
  int64_t offset = 0;
  int64_t reach = 0;
  uint32_t sp = 0, csp = 0;
  int pc;

  if (MODE == MODE_MATCH) {
    if (has_prefix) {
      int64_t idx = index_byte_dev(in, 0, l, (uint8_t)m.prefix);
      if (idx < 0) return 0;
      offset = idx;
    }
    if (memo) clear_visited(sc);
  } else if (MODE == MODE_FIND) {
    for (int i = 0; i < ncap; i++) caps[i] = 0;
    if (memo) clear_visited(sc);
  } else {
    offset = search_start;
    for (int i = 0; i < ncap; i++) caps[i] = CAP_ZERO;
    caps[0] = 0;
  }
  pc = m.start;
  if (MODE == MODE_FINDALL && resume_pc >= 0) {
    // resume an attempt after its leading class loop (kernels_btrun.cuh): the captures written before
    // the loop hold the attempt's start, the stack is empty (the loop's alternatives cannot succeed)
    pc = resume_pc;
    offset = resume_off;
    for (int i = 0; i < ncap; i++) if ((resume_start_caps >> i) & 1u) caps[i] = 0;
  }

  for (;;) {
    const uint4 I = insts[pc];
    const uint32_t op = I.x & 255u, ifl = I.x >> 8;
    bool fail = false;
    switch (op) {
      case OP_MATCH:
        if (MODE != MODE_MATCH) caps[1] = (int32_t)(offset - base);
        if (TRACK) *track_reach = reach;
        return 1;
      case OP_FAIL:
        if (MODE == MODE_FINDALL) { fail = true; break; }
        return 0;
      case OP_CAPTURE:
        if (MODE != MODE_MATCH) {
          if (per_capture) {
            if (sp >= sc.stack_cap) { atomicOr(err, ERR_STACK); return 0; }
            sc.stack[(size_t)sp * sc.stride + sc.tid] = make_uint2((uint32_t)caps[I.z], I.z | (2u << 16));
            sp++;
          }
          caps[I.z] = (int32_t)(offset - base);
        }
        pc = (int)I.y;
        break;
      case OP_NOP: case OP_ALTMATCH:
        pc = (int)I.y;
        break;
      case OP_RUNE1: {
        const uint32_t r = I.w;
        if (TRACK) reach = max(reach, offset + (r < 128 ? 1 : 4));   // (multi-byte literal: conservative)
        if (r < 128) {
          if (l <= offset || in[offset] != (uint8_t)r) { fail = true; break; }
          offset++;
        } else {
          uint8_t b[4]; int n;
          if (r < 0x800) { b[0] = 0xC0 | (r >> 6); b[1] = 0x80 | (r & 0x3F); n = 2; }
          else if (r < 0x10000) { b[0] = 0xE0 | (r >> 12); b[1] = 0x80 | ((r >> 6) & 0x3F); b[2] = 0x80 | (r & 0x3F); n = 3; }
          else { b[0] = 0xF0 | (r >> 18); b[1] = 0x80 | ((r >> 12) & 0x3F); b[2] = 0x80 | ((r >> 6) & 0x3F); b[3] = 0x80 | (r & 0x3F); n = 4; }
          if (l <= offset + n - 1) { fail = true; break; }
          for (int k = 0; k < n; k++) if (in[offset + k] != b[k]) fail = true;
          if (fail) break;
          offset += n;
        }
        pc = (int)I.y;
        break;
      }
      case OP_RUNE: {
        if (TRACK) reach = max(reach, offset + ((ifl & IF_UNICODE_CLASS) ? 4 : 1));
        if (l <= offset) { fail = true; break; }
        const uint32_t* bm = img + m.off_cls + 8 * pc;
        const uint32_t c = in[offset];
        if (!(ifl & IF_UNICODE_CLASS)) {
          if (!((bm[c >> 5] >> (c & 31)) & 1u)) { fail = true; break; }
          offset++;
        } else {
          const bool has_ascii = (bm[0] | bm[1] | bm[2] | bm[3]) != 0;
          if (has_ascii && c < 128) {
            if (!((bm[c >> 5] >> (c & 31)) & 1u)) { fail = true; break; }
            offset++;
          } else {
            int32_t r; const int width = decode_rune_dev(in + offset, l - offset, r);
            const uint32_t first = img[m.off_rng_idx + 2 * pc], cnt = img[m.off_rng_idx + 2 * pc + 1];
            bool found = false;
            for (uint32_t k = 0; k < cnt && !found; k++) {
              const int32_t lo = (int32_t)img[m.off_rng_pairs + 2 * (first + k)], hi = (int32_t)img[m.off_rng_pairs + 2 * (first + k) + 1];
              found = r >= lo && r <= hi;
            }
            if (!found) { fail = true; break; }
            offset += width;
          }
        }
        pc = (int)I.y;
        break;
      }
      case OP_ANY:
        if (TRACK) reach = max(reach, offset + 1);
        if (l <= offset) { fail = true; break; }
        offset++; pc = (int)I.y;
        break;
      case OP_ANYNOTNL:
        if (TRACK) reach = max(reach, offset + 1);
        if (l <= offset || in[offset] == '\n') { fail = true; break; }
        offset++; pc = (int)I.y;
        break;
      case OP_EMPTY: {
        const uint32_t a = I.z;
        if (TRACK) reach = max(reach, offset + 1);
        if ((a & EMPTY_BEGIN_TEXT) && offset != 0) fail = true;
        if ((a & EMPTY_END_TEXT) && offset != l) fail = true;
        if (!fail && (a & EMPTY_BEGIN_LINE) && offset != 0 && in[offset - 1] != '\n') fail = true;
        if (!fail && (a & EMPTY_END_LINE) && offset != l && in[offset] != '\n') fail = true;
        if (!fail && (a & (EMPTY_WORD | EMPTY_NOWORD))) {
          const bool pw = offset > 0 && is_word_byte(in[offset - 1]);
          const bool cw = offset < l && is_word_byte(in[offset]);
          if ((a & EMPTY_WORD) && pw == cw) fail = true;
          if ((a & EMPTY_NOWORD) && pw != cw) fail = true;
        }
        if (!fail) pc = (int)I.y;
        break;
      }
      case OP_ALT: {
        if (memo) {
          const int64_t idx = (int64_t)pc * (l + 1) + offset;
          const uint32_t word = (uint32_t)(idx >> 5), bit = 1u << (idx & 31);
          if (word >= sc.visited_words) { atomicOr(err, ERR_VISITED); return 0; }
          uint32_t* vp = &sc.visited[(size_t)word * sc.stride + sc.tid];
          if (*vp & bit) { fail = true; break; }
          *vp |= bit;
        }
        if (MODE == MODE_FINDALL && !memo && (ifl & IF_ATOMIC_LOOP) && sp > 0) {
          // Atomic loop (device_program.cu): the alternative stacked by the previous iteration (same exit,
          // one byte earlier, same captures) can only fail, so it is replaced instead of kept.
          uint2* top = &sc.stack[(size_t)(sp - 1) * sc.stride + sc.tid];
          const uint2 e = *top;
          if ((e.y & 0xFFFFu) == I.z && (int64_t)e.x + 1 == offset - base) {
            top->x = (uint32_t)(offset - base);
            pc = (int)I.y;
            break;
          }
        }
        if (sp >= sc.stack_cap) { atomicOr(err, ERR_STACK); return 0; }
        const int64_t rel = offset - base;
        if (rel > 0x7FFFFFFFll) { atomicOr(err, ERR_RANGE); return 0; }
        if (MODE != MODE_MATCH) {
          uint32_t ck = 0;
          if (!per_capture && (ifl & IF_ALT_CKPT)) {
            if (csp >= sc.cstack_cap) { atomicOr(err, ERR_CSTACK); return 0; }
            for (int k = 0; k < ncap; k++) sc.cstack[((size_t)csp * ncap + k) * sc.stride + sc.tid] = caps[k];
            csp++; ck = 1;
          }
          sc.stack[(size_t)sp * sc.stride + sc.tid] = make_uint2((uint32_t)rel, I.z | (ck << 16));
          sp++;
          pc = (int)I.y;
        } else if (ifl & IF_GREEDY_LOOP) {
          sc.stack[(size_t)sp * sc.stride + sc.tid] = make_uint2((uint32_t)rel, I.y);
          sp++;
          pc = (int)I.z;
        } else {
          sc.stack[(size_t)sp * sc.stride + sc.tid] = make_uint2((uint32_t)rel, I.z);
          sp++;
          pc = (int)I.y;
        }
        break;
      }
      default:
        return 0;
    }
    if (!fail) continue;

    // TryFallback
    if (needs_bt) {
      bool resumed = false;
      while (sp > 0) {
        sp--;
        const uint2 e = sc.stack[(size_t)sp * sc.stride + sc.tid];
        const uint32_t b = e.y & 0xFFFFu, t = e.y >> 16;
        if (MODE != MODE_MATCH && per_capture && t == 2) { caps[b] = (int32_t)e.x; continue; }
        offset = base + (int64_t)e.x;
        pc = (int)b;
        if (MODE != MODE_MATCH && !per_capture && t == 1 && csp > 0) {
          csp--;
          for (int k = 0; k < ncap; k++) caps[k] = sc.cstack[((size_t)csp * ncap + k) * sc.stride + sc.tid];
        }
        resumed = true;
        break;
      }
      if (resumed) continue;
    }
    if (MODE == MODE_FINDALL) {
      if (TRACK) { *track_fail = offset; *track_reach = reach; }
      return 0;
    }
    if (anchored) return 0;
    if (MODE == MODE_MATCH) {
      if (has_prefix) {
        offset++;
        if (l > offset) {
          const int64_t idx = index_byte_dev(in, offset, l, (uint8_t)m.prefix);
          if (idx < 0) return 0;
          offset = idx;
          if (memo && needs_bt) clear_visited(sc);
          pc = m.start;
          continue;
        }
        return 0;
      }
      if (l > offset) {
        pc = m.start; offset++;
        if (memo && needs_bt) clear_visited(sc);
        continue;
      }
      return 0;
    }
    // MODE_FIND
    if (l > offset) {
      offset++;
      for (int i = 0; i < ncap; i++) caps[i] = 0;
      csp = 0;
      if (memo) clear_visited(sc);
      caps[0] = (int32_t)offset;
      pc = m.start;
      continue;
    }
    return 0;
  }
}

// Offset record of a backtracking result (find.go:394-406): group i is set iff
// cap[2i] <= cap[2i+1] <= len(input); caps are base-relative with CAP_ZERO = absolute 0.
// `shift` is added to every absolute offset written (0, or the slice origin).
__device__ __forceinline__ void bt_emit_record(const int32_t* caps, int ncap, int64_t base, int64_t l, int64_t shift, int64_t* out) {
  for (int g = 0; g < ncap / 2; g++) {
    const int64_t a = caps[2 * g] == CAP_ZERO ? 0 : base + caps[2 * g];
    const int64_t b = caps[2 * g + 1] == CAP_ZERO ? 0 : base + caps[2 * g + 1];
    if (g == 0 || (a <= b && b <= l)) { out[2 * g] = a + shift; out[2 * g + 1] = b + shift; }
    else { out[2 * g] = -1; out[2 * g + 1] = -1; }
  }
}

// Thompson state-set MatchBytes (thompson.go:69-197).  Iterating the set bits of
// current & charStates visits the same states as the generated if-chain.
__device__ __forceinline__ uint64_t thompson_step(const DevMeta& m, const uint32_t* __restrict__ img, uint64_t cur, uint64_t char_mask, uint32_t c) {
  uint64_t next = 0, live = cur & char_mask;
  while (live) {
    const int s = __ffsll((long long)live) - 1;
    live &= live - 1;
    const uint32_t* cd = img + m.off_th_cond + 8 * s;
    if ((cd[c >> 5] >> (c & 31)) & 1u) next |= (uint64_t)img[m.off_th_eps + 2 * s] | ((uint64_t)img[m.off_th_eps + 2 * s + 1] << 32);
  }
  return next;
}

__device__ inline int thompson_match(const DevMeta& m, const uint32_t* __restrict__ img, const uint8_t* __restrict__ in, int64_t l) {
  const uint64_t start_closure = (uint64_t)m.th_start_lo | ((uint64_t)m.th_start_hi << 32);
  const uint64_t accept = (uint64_t)m.th_accept_lo | ((uint64_t)m.th_accept_hi << 32);
  const uint64_t char_mask = (uint64_t)m.th_char_lo | ((uint64_t)m.th_char_hi << 32);
  uint64_t cur;
  if (m.flags & F_ANCHORED) {
    cur = start_closure;
    for (int64_t off = 0; off < l; off++) {
      cur = thompson_step(m, img, cur, char_mask, in[off]);
      if (cur == 0) break;
      if (cur & accept) return 1;
    }
    return (cur & accept) != 0;
  }
  // The generated code restarts the simulation from every searchStart (thompson.go:69-131), O(l^2) when nothing
  // matches.  Its step is union-linear -- next = OR over the live states of a per-state, position-independent
  // closure mask -- so the union over all starts of "states alive at offset off" is one simulation that re-adds
  // the start closure at every offset, and "some start reaches an accepting state" is the same boolean.
  cur = start_closure;
  for (int64_t off = 0; off < l; off++) {
    if (cur & accept) return 1;
    cur = thompson_step(m, img, cur, char_mask, in[off]) | start_closure;
  }
  return (cur & accept) != 0;
}

// One start position of the TDFA walk (tdfa.go:905-994).  `in`/`l` is the slice FindBytes sees,
// `start` the slice-relative start (at_slice_start == (start == 0) there).  tags[]/mtags[] hold slice-relative positions, -1 = unset.
// Returns matchEnd (slice-relative) or -1; on a match mtags[] is the tag snapshot with group-0
// end forced and unset group ends defaulted to matchEnd (tdfa.go:1033-1046).
__device__ inline int64_t tdfa_walk(const DevMeta& m, const uint32_t* __restrict__ img, const uint8_t* __restrict__ in, int64_t l,
                                    int64_t start, bool at_slice_start, int64_t* mtags) {
  const int nt = m.t_ntags;
  int64_t tags[MAX_CAPS];
  for (int j = 0; j < nt; j++) { tags[j] = -1; mtags[j] = -1; }
  tags[0] = start;
  uint32_t state;
  if (at_slice_start) {  // `start == 0` in the generated code (tdfa.go:937)
    state = (uint32_t)m.t_start_begin;
    for (int k = 0; k < m.t_n_init_begin; k++) tags[img[m.off_t_init + k]] = start;
  } else {
    state = (uint32_t)m.t_start_any;
    for (int k = 0; k < m.t_n_init_any; k++) tags[img[m.off_t_init + m.t_n_init_begin + k]] = start;
  }
  int64_t match_end = -1;
  bool dirty = true;
  const uint32_t* trans = img + m.off_t_trans;
  const uint32_t* acc = img + m.off_t_accept;
  const uint32_t* aoff = img + m.off_t_alist_off;
  const uint32_t* alist = img + m.off_t_alist;
  {
    const uint32_t a = acc[state];
    if ((a & 1u) || (start == l && (a & 2u))) {
      match_end = start;
      for (int j = 0; j < nt; j++) mtags[j] = tags[j];
      dirty = false;
    }
  }
  for (int64_t i = start; i < l; i++) {
    const uint32_t c = in[i];
    if (c >= 128) break;
    const uint32_t e = trans[state * 128 + c];
    const uint32_t nx = e & 0xFFFFu;
    if (nx == TDFA_NONE) break;
    const uint32_t al = e >> 16;
    if (al) {
      for (uint32_t k = aoff[al]; k < aoff[al + 1]; k++) { const uint32_t x = alist[k]; tags[x & 0xFFFFu] = i + 1 - (int64_t)(x >> 16); }
      dirty = true;
    }
    state = nx;
    const uint32_t a = acc[state];
    if ((a & 1u) || (i == l - 1 && (a & 2u))) {
      const uint32_t aal = a >> 16;
      if (aal) {
        for (uint32_t k = aoff[aal]; k < aoff[aal + 1]; k++) { const uint32_t x = alist[k]; tags[x & 0xFFFFu] = i + 1 - (int64_t)(x >> 16); }
        dirty = true;
      }
      match_end = i + 1;
      if (dirty) { for (int j = 0; j < nt; j++) mtags[j] = tags[j]; dirty = false; }
    }
  }
  if (match_end >= 0) {
    mtags[1] = match_end;
    for (int g = 1; g < nt / 2; g++) if (mtags[2 * g] >= 0 && mtags[2 * g + 1] < 0) mtags[2 * g + 1] = match_end;
  }
  return match_end;
}

// findBytesInternal (tdfa.go:831-994): first start (prefix-accelerated) whose walk matches.
// out[]: offset record relative to the slice, shifted by `shift`.
__device__ inline int tdfa_find(const DevMeta& m, const uint32_t* __restrict__ img, const uint8_t* __restrict__ in, int64_t l,
                                int64_t shift, int64_t* out) {
  const bool has_prefix = (m.flags & F_HAS_PREFIX) && !(m.flags & F_ANCHORED);
  int64_t mtags[MAX_CAPS];
  for (int64_t start = 0; start <= l; start++) {
    if (has_prefix) {
      const int64_t idx = index_byte_dev(in, start, l, (uint8_t)m.prefix);
      if (idx < 0) break;
      start = idx;
    }
    if (tdfa_walk(m, img, in, l, start, start == 0, mtags) >= 0) {
      const int nt = m.t_ntags;
      out[0] = mtags[0] + shift; out[1] = mtags[1] + shift;
      for (int g = 1; g < nt / 2; g++) {
        if (mtags[2 * g] >= 0) { out[2 * g] = mtags[2 * g] + shift; out[2 * g + 1] = mtags[2 * g + 1] + shift; }
        else { out[2 * g] = -1; out[2 * g + 1] = -1; }
      }
      return 1;
    }
  }
  return 0;
}

}  // namespace rgx
