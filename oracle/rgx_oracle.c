/*
 * rgx_oracle.c -- CPU ORACLE for the regengo matching hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is a plain-C restatement of the loops that KromDaniel/regengo GENERATES per pattern
 * (the reference has no runtime matcher; the generated Go file is the matcher).  It is used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the
 * checker and the CPU baseline.  The product (regengo_b200/, libregengo_b200.so) never links,
 * imports or calls it.
 *
 * Pinning (how this oracle is tied to the reference, since Go is not installed here and the
 * generated code cannot be executed):
 *   - tests/test_oracle_goldens.py runs it on program blobs built DIRECTLY from the programs
 *     mined out of the reference's 24 checked-in generated Go files (tests/golden/
 *     generated_goldens.json: instruction listings, class bitmaps, Thompson masks, TDFA tables)
 *     and checks every curated input of scripts/curated/cases.go against leftmost-first
 *     semantics (the assertion the reference's own generated tests make against Go's regexp);
 *   - tests/test_oracle_corpus.py does the same for the 238-entry tests/e2e/testdata.json corpus;
 *   - tests/test_oracle_quirks.py holds one witness per reference quirk and the
 *     stream known-answer tests (tests/integration/streaming/streaming_test.go:190-316).
 *
 * Each function cites the reference file:line it follows (paths under /root/reference).
 * Input: the program blob (layout in regengo_b200/csrc/blob.hpp; this file has its own reader).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define _GNU_SOURCE_MEMMEM 1
void* memmem(const void*, size_t, const void*, size_t);

/* ---- blob reader (independent of the product's) ---------------------------------------------- */
enum { H_MAGIC = 0, H_VERSION, H_WORDS, H_NINST, H_START, H_NUMCAP, H_FLAGS, H_PREFIX, H_MATCH_ENGINE,
       H_FIND_ENGINE, H_MINLEN, H_MAXLEN, H_LEFTOVER, H_MINBUF, H_OFF_INST, H_OFF_CLASS, H_OFF_RANGES,
       H_NRANGE_PAIRS, H_OFF_THOMPSON, H_OFF_TDFA, H_OFF_NAMES, H_NAMES_WORDS };
enum { F_ANCHORED = 1, F_NEEDS_BT = 2, F_HAS_PREFIX = 4, F_MATCH_MEMO = 8, F_FIND_MEMO = 16, F_PER_CAPTURE = 32,
       F_HAS_CAPTURES = 64 };
enum { IF_ALT_CKPT = 1, IF_GREEDY_LOOP = 2, IF_UNICODE_CLASS = 4, IF_CHAR_STATE = 8 };
enum { OP_ALT = 0, OP_ALTMATCH, OP_CAPTURE, OP_EMPTY, OP_MATCH, OP_FAIL, OP_NOP, OP_RUNE, OP_RUNE1, OP_ANY, OP_ANYNOTNL };
enum { EMPTY_BEGIN_LINE = 1, EMPTY_END_LINE = 2, EMPTY_BEGIN_TEXT = 4, EMPTY_END_TEXT = 8, EMPTY_WORD = 16, EMPTY_NOWORD = 32 };
enum { T_NSTATES = 0, T_NTAGS, T_START_BEGIN, T_START_ANY, T_NINIT_BEGIN, T_NINIT_ANY, T_NACTS, T_NACC_ACTS,
       T_MAX_ACTS, T_MAX_ACC_ACTS, T_WORDS_HDR = 12 };

typedef struct {
  const uint32_t* w;
  int n_inst, start, num_cap;
  uint32_t flags;
  int prefix, match_engine, find_engine;
  const uint32_t *inst, *cls, *rng_idx, *rng_pairs, *th, *th_eps, *th_cond;
  /* tdfa */
  int t_ns, t_ntags, t_start_begin, t_start_any, t_nib, t_nia;
  const uint32_t *t_init_begin, *t_init_any, *t_accept, *t_accept_eot, *t_act_off, *t_acts, *t_acc_off, *t_acc_acts;
  const int32_t* t_trans;
  /* scratch, grown on demand */
  int64_t* stack; size_t stack_cap;      /* 3 ints per entry */
  int64_t* cstack; size_t cstack_cap;
  uint32_t* visited; size_t visited_cap;
} orc_prog;

orc_prog* orc_open(const uint32_t* w, size_t n_words) {
  if (n_words < 32 || w[H_MAGIC] != 0x42584752u || w[H_WORDS] != n_words) return NULL;
  orc_prog* P = (orc_prog*)calloc(1, sizeof(orc_prog));
  P->w = w;
  P->n_inst = (int)w[H_NINST]; P->start = (int)w[H_START]; P->num_cap = (int)w[H_NUMCAP];
  P->flags = w[H_FLAGS]; P->prefix = (int)w[H_PREFIX];
  P->match_engine = (int)w[H_MATCH_ENGINE]; P->find_engine = (int)w[H_FIND_ENGINE];
  P->inst = w + w[H_OFF_INST]; P->cls = w + w[H_OFF_CLASS];
  P->rng_idx = w + w[H_OFF_RANGES]; P->rng_pairs = P->rng_idx + 2 * (size_t)P->n_inst;
  P->th = w + w[H_OFF_THOMPSON]; P->th_eps = P->th + 4; P->th_cond = P->th_eps + 2 * (size_t)P->n_inst;
  if (w[H_OFF_TDFA]) {
    const uint32_t* t = w + w[H_OFF_TDFA];
    P->t_ns = (int)t[T_NSTATES]; P->t_ntags = (int)t[T_NTAGS];
    P->t_start_begin = (int)t[T_START_BEGIN]; P->t_start_any = (int)t[T_START_ANY];
    P->t_nib = (int)t[T_NINIT_BEGIN]; P->t_nia = (int)t[T_NINIT_ANY];
    const uint32_t* p = t + T_WORDS_HDR;
    P->t_init_begin = p; p += P->t_nib;
    P->t_init_any = p; p += P->t_nia;
    P->t_trans = (const int32_t*)p; p += (size_t)P->t_ns * 128;
    P->t_accept = p; p += P->t_ns;
    P->t_accept_eot = p; p += P->t_ns;
    P->t_act_off = p; p += (size_t)P->t_ns * 128 + 1;
    P->t_acts = p; p += 2 * (size_t)t[T_NACTS];
    P->t_acc_off = p; p += P->t_ns + 1;
    P->t_acc_acts = p;
  }
  return P;
}
void orc_close(orc_prog* P) { if (!P) return; free(P->stack); free(P->cstack); free(P->visited); free(P); }
int orc_num_cap(const orc_prog* P) { return P->num_cap; }

/* ---- helpers ------------------------------------------------------------------------------------ */
static inline int is_word(uint8_t b) { /* compiler.go:674-688 isWordChar */
  return (b >= 'a' && b <= 'z') || (b >= 'A' && b <= 'Z') || (b >= '0' && b <= '9') || b == '_';
}
/* Go unicode/utf8.DecodeRune [go-stdlib]: (RuneError=0xFFFD, 1) on any invalid/short encoding. */
static int decode_rune(const uint8_t* p, int64_t n, int32_t* r) {
  if (n < 1) { *r = 0xFFFD; return 0; }
  uint8_t b0 = p[0];
  if (b0 < 0x80) { *r = b0; return 1; }
  if (b0 < 0xC2) { *r = 0xFFFD; return 1; }
  if (b0 < 0xE0) {
    if (n < 2 || (p[1] & 0xC0) != 0x80) { *r = 0xFFFD; return 1; }
    *r = ((b0 & 0x1F) << 6) | (p[1] & 0x3F); return 2;
  }
  if (b0 < 0xF0) {
    if (n < 3) { *r = 0xFFFD; return 1; }
    uint8_t lo = 0x80, hi = 0xBF;
    if (b0 == 0xE0) lo = 0xA0; else if (b0 == 0xED) hi = 0x9F;
    if (p[1] < lo || p[1] > hi || (p[2] & 0xC0) != 0x80) { *r = 0xFFFD; return 1; }
    *r = ((b0 & 0x0F) << 12) | ((p[1] & 0x3F) << 6) | (p[2] & 0x3F); return 3;
  }
  if (b0 < 0xF5) {
    if (n < 4) { *r = 0xFFFD; return 1; }
    uint8_t lo = 0x80, hi = 0xBF;
    if (b0 == 0xF0) lo = 0x90; else if (b0 == 0xF4) hi = 0x8F;
    if (p[1] < lo || p[1] > hi || (p[2] & 0xC0) != 0x80 || (p[3] & 0xC0) != 0x80) { *r = 0xFFFD; return 1; }
    *r = ((b0 & 0x07) << 18) | ((p[1] & 0x3F) << 12) | ((p[2] & 0x3F) << 6) | (p[3] & 0x3F); return 4;
  }
  *r = 0xFFFD; return 1;
}
static int encode_rune(uint32_t r, uint8_t* b) { /* utf8.EncodeRune [go-stdlib] */
  if (r < 0x80) { b[0] = (uint8_t)r; return 1; }
  if (r < 0x800) { b[0] = 0xC0 | (r >> 6); b[1] = 0x80 | (r & 0x3F); return 2; }
  if (r > 0x10FFFF || (r >= 0xD800 && r <= 0xDFFF)) r = 0xFFFD;
  if (r < 0x10000) { b[0] = 0xE0 | (r >> 12); b[1] = 0x80 | ((r >> 6) & 0x3F); b[2] = 0x80 | (r & 0x3F); return 3; }
  b[0] = 0xF0 | (r >> 18); b[1] = 0x80 | ((r >> 12) & 0x3F); b[2] = 0x80 | ((r >> 6) & 0x3F); b[3] = 0x80 | (r & 0x3F); return 4;
}
static int64_t index_byte(const uint8_t* p, int64_t n, int b) { /* bytes.IndexByte */
  if (n <= 0) return -1;
  const uint8_t* q = (const uint8_t*)memchr(p, b, (size_t)n);
  return q ? (int64_t)(q - p) : -1;
}
static void grow_stack(orc_prog* P, size_t need) {
  if (need <= P->stack_cap) return;
  size_t c = P->stack_cap ? P->stack_cap * 2 : 96;
  while (c < need) c *= 2;
  P->stack = (int64_t*)realloc(P->stack, c * sizeof(int64_t)); P->stack_cap = c;
}
static void grow_cstack(orc_prog* P, size_t need) {
  if (need <= P->cstack_cap) return;
  size_t c = P->cstack_cap ? P->cstack_cap * 2 : 256;
  while (c < need) c *= 2;
  P->cstack = (int64_t*)realloc(P->cstack, c * sizeof(int64_t)); P->cstack_cap = c;
}
static size_t visited_words(const orc_prog* P, int64_t l) { return (size_t)(((int64_t)P->n_inst * (l + 1) + 31) / 32); }
static void visited_alloc(orc_prog* P, int64_t l) {
  size_t n = visited_words(P, l);
  if (n > P->visited_cap) { P->visited = (uint32_t*)realloc(P->visited, n * sizeof(uint32_t)); P->visited_cap = n; }
  memset(P->visited, 0, n * sizeof(uint32_t));
}

/* ---- the backtracking goto-machine --------------------------------------------------------------
 * One interpreter for the three generated variants:
 *   MODE_MATCH   MatchBytes          compiler.go:740-871 + backtracking.go:9-77
 *   MODE_FIND    FindBytesReuse      find.go:469-591     + backtracking.go:83-165
 *   MODE_FINDALL FindAllBytesAppend  find.go:130-316 (one attempt at searchStart; caller iterates)
 * Instruction bodies: instructions.go:51-605, captures.go:123-158.
 * Returns 1 on Match (caps filled for FIND modes), 0 otherwise.  For MODE_FINDALL a 0 means
 * "searchStart++".                                                                              */
enum { MODE_MATCH = 0, MODE_FIND = 1, MODE_FINDALL = 2 };

static int bt_machine(orc_prog* P, int mode, const uint8_t* in, int64_t l, int64_t search_start, int64_t* caps) {
  const int ncap = P->num_cap;
  const int anchored = (P->flags & F_ANCHORED) != 0;
  const int needs_bt = (P->flags & F_NEEDS_BT) != 0;
  const int per_capture = (P->flags & F_PER_CAPTURE) != 0;
  const int memo = mode == MODE_MATCH ? (P->flags & F_MATCH_MEMO) != 0 : (P->flags & F_FIND_MEMO) != 0;
  const int has_prefix = mode == MODE_MATCH && (P->flags & F_HAS_PREFIX) && !anchored;
  int64_t offset = 0;
  size_t sp = 0, csp = 0; /* stack entries *3, capture stack ints */
  int pc;

  if (mode == MODE_MATCH) {
    if (has_prefix) { /* compiler.go:751-764 */
      int64_t idx = index_byte(in, l, P->prefix);
      if (idx < 0) return 0;
      offset = idx;
    }
    if (memo) visited_alloc(P, l);
  } else if (mode == MODE_FIND) {
    for (int i = 0; i < ncap; i++) caps[i] = 0; /* var captures [N]int; captures[0] = 0 (find.go:485,523) */
    if (memo) visited_alloc(P, l);
  } else {
    offset = search_start; /* find.go:213-230; visited is NOT cleared here */
    for (int i = 0; i < ncap; i++) caps[i] = 0;
    caps[0] = search_start;
  }
  pc = P->start;

  for (;;) {
    const uint32_t* I = P->inst + 4 * (size_t)pc;
    const int op = (int)(I[0] & 255);
    const uint32_t ifl = I[0] >> 8;
    switch (op) {
      case OP_MATCH:
        if (mode != MODE_MATCH) caps[1] = offset; /* find.go:386,680 */
        return 1;
      case OP_FAIL:
        if (mode == MODE_FINDALL) goto fallback; /* find.go:329-335 */
        return 0;                                /* instructions.go:62-66, find.go:723-727: plain return */
      case OP_CAPTURE:
        if (mode != MODE_MATCH) {
          if (per_capture) { /* captures.go:129-146 */
            grow_stack(P, sp + 3);
            P->stack[sp] = caps[I[2]]; P->stack[sp + 1] = I[2]; P->stack[sp + 2] = 2; sp += 3;
          }
          caps[I[2]] = offset;
        }
        pc = (int)I[1];
        continue;
      case OP_NOP: case OP_ALTMATCH:
        pc = (int)I[1];
        continue;
      case OP_RUNE1: { /* instructions.go:106-175 */
        uint32_t r = I[3];
        if (r < 128) {
          if (l <= offset) goto fallback;
          if (in[offset] != (uint8_t)r) goto fallback;
          offset++;
        } else {
          uint8_t b[4]; int n = encode_rune(r, b);
          if (l <= offset + n - 1) goto fallback;
          for (int k = 0; k < n; k++) if (in[offset + k] != b[k]) goto fallback;
          offset += n;
        }
        pc = (int)I[1];
        continue;
      }
      case OP_RUNE: { /* instructions.go:178-295 */
        if (l <= offset) goto fallback;
        const uint32_t* bm = P->cls + 8 * (size_t)pc;
        uint8_t c = in[offset];
        if (!(ifl & IF_UNICODE_CLASS)) {
          if (!((bm[c >> 5] >> (c & 31)) & 1)) goto fallback;
          offset++;
        } else {
          int has_ascii = 0;
          for (int k = 0; k < 4; k++) if (bm[k]) has_ascii = 1;
          if (has_ascii && c < 128) {
            if (!((bm[c >> 5] >> (c & 31)) & 1)) goto fallback;
            offset++;
          } else {
            int32_t r; int width = decode_rune(in + offset, l - offset, &r);
            uint32_t first = P->rng_idx[2 * (size_t)pc], cnt = P->rng_idx[2 * (size_t)pc + 1];
            int found = 0;
            for (uint32_t k = 0; k < cnt; k++) {
              int32_t lo = (int32_t)P->rng_pairs[2 * (first + k)], hi = (int32_t)P->rng_pairs[2 * (first + k) + 1];
              if (r >= lo && r <= hi) { found = 1; break; }
            }
            if (!found) goto fallback;
            offset += width;
          }
        }
        pc = (int)I[1];
        continue;
      }
      case OP_ANY: /* instructions.go:298-309 */
        if (l <= offset) goto fallback;
        offset++; pc = (int)I[1];
        continue;
      case OP_ANYNOTNL: /* instructions.go:312-328 */
        if (l <= offset || in[offset] == '\n') goto fallback;
        offset++; pc = (int)I[1];
        continue;
      case OP_EMPTY: { /* instructions.go:492-595 */
        uint32_t a = I[2];
        if ((a & EMPTY_BEGIN_TEXT) && offset != 0) goto fallback;
        if ((a & EMPTY_END_TEXT) && offset != l) goto fallback;
        if ((a & EMPTY_BEGIN_LINE) && offset != 0 && in[offset - 1] != '\n') goto fallback;
        if ((a & EMPTY_END_LINE) && offset != l && in[offset] != '\n') goto fallback;
        if (a & (EMPTY_WORD | EMPTY_NOWORD)) {
          int pw = offset > 0 && is_word(in[offset - 1]);
          int cw = offset < l && is_word(in[offset]);
          if ((a & EMPTY_WORD) && pw == cw) goto fallback;
          if ((a & EMPTY_NOWORD) && pw != cw) goto fallback;
        }
        pc = (int)I[1];
        continue;
      }
      case OP_ALT: { /* instructions.go:331-457 */
        if (memo) {
          int64_t idx = (int64_t)pc * (l + 1) + offset;
          uint32_t bit = 1u << (idx & 31);
          if (P->visited[idx >> 5] & bit) goto fallback;
          P->visited[idx >> 5] |= bit;
        }
        grow_stack(P, sp + 3);
        if (mode != MODE_MATCH) {
          int ck = 0;
          if (!per_capture && (ifl & IF_ALT_CKPT)) {
            grow_cstack(P, csp + ncap);
            memcpy(P->cstack + csp, caps, sizeof(int64_t) * ncap); csp += ncap; ck = 1;
          }
          P->stack[sp] = offset; P->stack[sp + 1] = I[2]; P->stack[sp + 2] = ck; sp += 3;
          pc = (int)I[1];
        } else if (ifl & IF_GREEDY_LOOP) { /* continuation first, loop body on the stack */
          P->stack[sp] = offset; P->stack[sp + 1] = I[1]; P->stack[sp + 2] = 0; sp += 3;
          pc = (int)I[2];
        } else {
          P->stack[sp] = offset; P->stack[sp + 1] = I[2]; P->stack[sp + 2] = 0; sp += 3;
          pc = (int)I[1];
        }
        continue;
      }
      default:
        return 0;
    }
  fallback:
    /* TryFallback: backtracking.go:9-77 (Match), :83-165 (Find), find.go:232-289 (FindAll) */
    if (needs_bt) {
      int resumed = 0;
      while (sp > 0) {
        sp -= 3;
        int64_t a = P->stack[sp], b = P->stack[sp + 1], t = P->stack[sp + 2];
        if (mode != MODE_MATCH && per_capture && t == 2) { caps[b] = a; continue; }
        offset = a; pc = (int)b;
        if (mode != MODE_MATCH && !per_capture && t == 1 && csp > 0) {
          csp -= ncap; memcpy(caps, P->cstack + csp, sizeof(int64_t) * ncap);
        }
        resumed = 1;
        break;
      }
      if (resumed) continue;
    }
    /* stack empty: restart (Match/Find) or give up on this searchStart (FindAll) */
    if (mode == MODE_FINDALL) return 0;
    if (anchored) return 0;
    if (mode == MODE_MATCH) {
      if (has_prefix) { /* compiler.go:816-841 / backtracking.go:24-43 */
        offset++;
        if (l > offset) {
          int64_t idx = index_byte(in + offset, l - offset, P->prefix);
          if (idx < 0) return 0;
          offset += idx;
          if (memo && needs_bt) memset(P->visited, 0, visited_words(P, l) * sizeof(uint32_t));
          pc = P->start;
          continue;
        }
        return 0;
      }
      if (l > offset) { /* restart at FAILURE offset + 1 */
        pc = P->start; offset++;
        if (memo && needs_bt) memset(P->visited, 0, visited_words(P, l) * sizeof(uint32_t));
        continue;
      }
      return 0;
    }
    /* MODE_FIND: backtracking.go:94-118 / find.go:546-572 */
    if (l > offset) {
      offset++;
      for (int i = 0; i < ncap; i++) caps[i] = 0;
      csp = 0;
      if (memo) memset(P->visited, 0, visited_words(P, l) * sizeof(uint32_t));
      caps[0] = offset;
      pc = P->start;
      continue;
    }
    return 0;
  }
}

/* group i of a backtracking result (find.go:394-406): set iff cap[2i] <= cap[2i+1] <= len(input) */
static void bt_emit(const int64_t* caps, int ncap, int64_t l, int64_t base, int64_t* out) {
  out[0] = base + caps[0]; out[1] = base + caps[1];
  for (int g = 1; g < ncap / 2; g++) {
    if (caps[2 * g] <= caps[2 * g + 1] && caps[2 * g + 1] <= l) { out[2 * g] = base + caps[2 * g]; out[2 * g + 1] = base + caps[2 * g + 1]; }
    else { out[2 * g] = -1; out[2 * g + 1] = -1; }
  }
}

/* ---- Thompson bitset MatchBytes: thompson.go:69-197 ------------------------------------------------ */
static int thompson_match(const orc_prog* P, const uint8_t* in, int64_t l) {
  const uint64_t start_closure = (uint64_t)P->th[0] | ((uint64_t)P->th[1] << 32);
  const uint64_t accept = (uint64_t)P->th[2] | ((uint64_t)P->th[3] << 32);
  const int n = P->n_inst < 64 ? P->n_inst : 64;
  uint64_t cur;
#define TH_STEP(c)                                                                                   \
  do {                                                                                               \
    uint64_t next = 0;                                                                               \
    for (int s = 0; s < n; s++) {                                                                    \
      if (!((P->inst[4 * s] >> 8) & IF_CHAR_STATE)) continue;                                        \
      if ((cur >> s) & 1) {                                                                          \
        const uint32_t* cd = P->th_cond + 8 * (size_t)s;                                             \
        if ((cd[(c) >> 5] >> ((c) & 31)) & 1) next |= (uint64_t)P->th_eps[2 * s] | ((uint64_t)P->th_eps[2 * s + 1] << 32); \
      }                                                                                              \
    }                                                                                                \
    cur = next;                                                                                      \
  } while (0)
  if (P->flags & F_ANCHORED) { /* thompson.go:94-107 */
    cur = start_closure;
    for (int64_t off = 0; off < l; off++) {
      uint8_t c = in[off];
      TH_STEP(c);
      if (cur == 0) break;
      if (cur & accept) return 1;
    }
    return (cur & accept) != 0;
  }
  for (int64_t ss = 0; ss <= l; ss++) { /* thompson.go:109-128 */
    cur = start_closure;
    if (cur & accept) return 1;
    for (int64_t off = ss; off < l; off++) {
      uint8_t c = in[off];
      TH_STEP(c);
      if (cur == 0) break;
      if (cur & accept) return 1;
    }
  }
  return 0;
}

/* ---- TDFA findBytesInternal: tdfa.go:831-1052 ------------------------------------------------------
 * tags_out[ntags]: final tags (group 0 end forced, unset end := match end); -1 = unset.          */
#define MAX_TAGS 256
static int tdfa_find(const orc_prog* P, const uint8_t* in, int64_t l, int64_t* tags_out) {
  const int nt = P->t_ntags;
  int64_t tags[MAX_TAGS], mtags[MAX_TAGS];
  int64_t match_end = -1;
  const int has_prefix = (P->flags & F_HAS_PREFIX) && !(P->flags & F_ANCHORED);
  for (int j = 0; j < nt; j++) mtags[j] = -1;
  for (int64_t start = 0; start <= l; start++) {
    if (has_prefix) {
      int64_t idx = index_byte(in + start, l - start, P->prefix);
      if (idx < 0) break;
      start += idx;
    }
    for (int j = 0; j < nt; j++) tags[j] = -1;
    tags[0] = start;
    int state;
    if (start == 0) { state = P->t_start_begin; for (int k = 0; k < P->t_nib; k++) tags[P->t_init_begin[k]] = start; }
    else { state = P->t_start_any; for (int k = 0; k < P->t_nia; k++) tags[P->t_init_any[k]] = start; }
    if (P->t_accept[state]) { match_end = start; memcpy(mtags, tags, sizeof(int64_t) * nt); }
    if (start == l && P->t_accept_eot[state]) { match_end = start; memcpy(mtags, tags, sizeof(int64_t) * nt); }
    for (int64_t i = start; i < l; i++) {
      uint8_t c = in[i];
      if (c >= 128) break;
      int32_t nx = P->t_trans[(size_t)state * 128 + c];
      if (nx < 0) break;
      size_t e = (size_t)state * 128 + c;
      for (uint32_t a = P->t_act_off[e]; a < P->t_act_off[e + 1]; a++)
        tags[P->t_acts[2 * a]] = i + 1 - (int64_t)P->t_acts[2 * a + 1];
      state = nx;
      if (P->t_accept[state]) {
        for (uint32_t a = P->t_acc_off[state]; a < P->t_acc_off[state + 1]; a++)
          tags[P->t_acc_acts[2 * a]] = i + 1 - (int64_t)P->t_acc_acts[2 * a + 1];
        match_end = i + 1; memcpy(mtags, tags, sizeof(int64_t) * nt);
      }
      if (i == l - 1 && P->t_accept_eot[state]) {
        for (uint32_t a = P->t_acc_off[state]; a < P->t_acc_off[state + 1]; a++)
          tags[P->t_acc_acts[2 * a]] = i + 1 - (int64_t)P->t_acc_acts[2 * a + 1];
        match_end = i + 1; memcpy(mtags, tags, sizeof(int64_t) * nt);
      }
    }
    if (match_end >= 0) {
      mtags[1] = match_end;
      for (int g = 1; g < nt / 2; g++) /* tdfa.go:1033-1046 */
        if (mtags[2 * g] >= 0 && mtags[2 * g + 1] < 0) mtags[2 * g + 1] = mtags[1];
      memcpy(tags_out, mtags, sizeof(int64_t) * nt);
      return 1;
    }
  }
  return 0;
}
static void tdfa_emit(const int64_t* tags, int nt, int64_t base, int64_t* out) {
  out[0] = base + tags[0]; out[1] = base + tags[1];
  for (int g = 1; g < nt / 2; g++) {
    if (tags[2 * g] >= 0) { out[2 * g] = base + tags[2 * g]; out[2 * g + 1] = base + tags[2 * g + 1]; }
    else { out[2 * g] = -1; out[2 * g + 1] = -1; }
  }
}

/* ---- public: one input ------------------------------------------------------------------------------ */
int orc_match(orc_prog* P, const uint8_t* in, int64_t l) {
  if (P->match_engine == 1) return thompson_match(P, in, l);
  return bt_machine(P, MODE_MATCH, in, l, 0, NULL);
}

/* FindBytes: out[num_cap] offset record relative to `in`; returns found. */
int orc_find(orc_prog* P, const uint8_t* in, int64_t l, int64_t* out) {
  if (P->find_engine == 2) {
    int64_t tags[MAX_TAGS];
    if (!tdfa_find(P, in, l, tags)) return 0;
    tdfa_emit(tags, P->t_ntags, 0, out);
    return 1;
  }
  if (P->find_engine != 1) return -1;
  int64_t caps[MAX_TAGS];
  if (!bt_machine(P, MODE_FIND, in, l, 0, caps)) return 0;
  bt_emit(caps, P->num_cap, l, 0, out);
  return 1;
}

/* FindAllBytes(input, n): writes up to cap records, returns the reference's len(result). */
int64_t orc_find_all(orc_prog* P, const uint8_t* in, int64_t l, int64_t n_limit, int64_t* out, int64_t cap) {
  const int nc = P->num_cap;
  int64_t count = 0;
  if (n_limit == 0) return 0;
  if (P->find_engine == 2) { /* compiler.go:602-655 (stride rule) */
    int64_t offset = 0, tags[MAX_TAGS];
    while (offset < l) {
      if (!tdfa_find(P, in + offset, l - offset, tags)) break;
      if (count < cap) tdfa_emit(tags, P->t_ntags, offset, out + count * nc);
      count++;
      if (n_limit > 0 && count >= n_limit) break;
      int64_t mlen = tags[1] - tags[0];
      if (mlen > 0) offset += mlen; else offset++;
    }
    return count;
  }
  if (P->find_engine != 1) return -1;
  /* find.go:130-316 */
  int64_t caps[MAX_TAGS], ss = 0;
  const int anchored = (P->flags & F_ANCHORED) != 0;
  if (P->flags & F_FIND_MEMO) visited_alloc(P, l);
  for (;;) {
    if (n_limit > 0 && count >= n_limit) break;
    if (anchored && ss > 0) break;
    if (ss >= l) break;
    if (bt_machine(P, MODE_FINDALL, in, l, ss, caps)) {
      if (count < cap) bt_emit(caps, nc, l, 0, out + count * nc);
      count++;
      if (caps[1] > ss) ss = caps[1]; else ss++;
    } else {
      ss++;
    }
  }
  return count;
}

/* stream.Config Validate + ApplyDefaults: stream/stream.go:96-134 */
int orc_stream_config(const orc_prog* P, int64_t buffer_size, int64_t max_leftover, int64_t* eb, int64_t* el) {
  int64_t min_buffer = (int32_t)P->w[H_MINBUF], def_left = (int32_t)P->w[H_LEFTOVER];
  if (buffer_size > 0 && buffer_size < min_buffer) return -6;
  if (buffer_size == 0) buffer_size = 64 * 1024;
  if (buffer_size < min_buffer) buffer_size = min_buffer;
  if (max_leftover == 0) max_leftover = def_left;
  int64_t max_allowed = buffer_size / 2;
  if (max_leftover != -1 && max_leftover > max_allowed) max_leftover = max_allowed;
  *eb = buffer_size; *el = max_leftover;
  return 0;
}

/* FindReader over a reader that fills every Read (bytes.Reader): streaming.go:85-255.
 * Per match: StreamOffset, ChunkIndex and the offset record as ABSOLUTE stream offsets of the
 * result slices.  Unset TDFA groups are reported (-1,-1) (the reference leaves the reused struct
 * field stale there).  Literal emulation: it really carries `leftover` around.       */
int64_t orc_find_reader(orc_prog* P, const uint8_t* stream, int64_t len, int64_t buffer_size, int64_t max_leftover,
                        int64_t* out_soff, int32_t* out_chunk, int64_t* out, int64_t cap) {
  int64_t B, L;
  int rc = orc_stream_config(P, buffer_size, max_leftover, &B, &L);
  if (rc < 0) return rc;
  const int nc = P->num_cap;
  uint8_t* buf = (uint8_t*)malloc((size_t)B);
  int64_t leftover = 0, stream_offset = 0, rd = 0, count = 0;
  int32_t chunk_index = 0;
  int64_t rec[MAX_TAGS];
  for (;;) {
    int64_t want = B - leftover;
    int64_t n = len - rd < want ? len - rd : want;
    int eof_now = (n == 0); /* bytes.Reader: (0, io.EOF) once drained; err==nil otherwise */
    if (n > 0) { memcpy(buf + leftover, stream + rd, (size_t)n); rd += n; }
    int flush = 0;
    int64_t data_len; int is_full;
    if (eof_now) { /* streaming.go:128-172 */
      if (leftover <= 0) break;
      flush = 1; data_len = leftover; is_full = 0;
    } else {
      data_len = leftover + n; is_full = (n == B - leftover);
    }
    int64_t search_pos = 0, committed = 0;
    while (search_pos < data_len) {
      if (orc_find(P, buf + search_pos, data_len - search_pos, rec) != 1) break;
      int64_t mlen = rec[1] - rec[0];
      /* bytes.Index(chunk[searchPos:], result.Match)   */
      int64_t midx;
      if (mlen == 0) midx = 0;
      else {
        const uint8_t* q = (const uint8_t*)memmem(buf + search_pos, (size_t)(data_len - search_pos), buf + search_pos + rec[0], (size_t)mlen);
        if (!q) break;
        midx = (int64_t)(q - (buf + search_pos));
      }
      int64_t mstart = search_pos + midx, mend = mstart + mlen;
      if (!flush && is_full && mend > data_len - L) break;
      if (count < cap) {
        out_soff[count] = stream_offset + mstart;
        out_chunk[count] = chunk_index;
        for (int k = 0; k < nc; k++) out[count * nc + k] = rec[k] < 0 ? -1 : stream_offset + search_pos + rec[k];
      }
      count++;
      committed = mend;
      if (mlen > 0) search_pos = mend; else search_pos++;
    }
    if (flush) break;
    if (is_full) {
      int64_t keep_from = data_len - L;
      if (keep_from < committed) keep_from = committed;
      leftover = data_len - keep_from;
      stream_offset += keep_from;
      memmove(buf, buf + keep_from, (size_t)leftover);
    } else {
      leftover = 0;
    }
    chunk_index++;
  }
  free(buf);
  return count;
}

/* ---- batch helpers (CPU baseline timing; one orc_prog per thread is the caller's job) -------------- */
void orc_match_batch(orc_prog* P, const uint8_t* bytes, const uint64_t* offs, uint64_t n, uint8_t* out) {
  for (uint64_t i = 0; i < n; i++) out[i] = (uint8_t)orc_match(P, bytes + offs[i], (int64_t)(offs[i + 1] - offs[i]));
}
void orc_find_batch(orc_prog* P, const uint8_t* bytes, const uint64_t* offs, uint64_t n, uint8_t* found, int64_t* out) {
  for (uint64_t i = 0; i < n; i++) {
    int r = orc_find(P, bytes + offs[i], (int64_t)(offs[i + 1] - offs[i]), out + i * P->num_cap);
    found[i] = (uint8_t)(r == 1);
  }
}

/* ---- ReplaceAllBytesAppend (internal/compiler/replace.go:192-273) ------------------------------------
 * Template: replace.Parse (replace/template.go:60-163), restated bytewise as the Go code indexes its string.
 * isNameStart / isNameContinue take rune(byte): a byte >= 0x80 is judged as the Latin-1 code point of that value
 * (unicode.IsLetter over U+0080..U+00FF: AA, B5, BA, C0-D6, D8-F6, F8-FF; unicode.IsDigit: none).  Braced names
 * are checked rune by rune in the reference; this restatement only decides ASCII content and reports -8 for a
 * braced reference with a byte >= 0x80 (tests keep to what it decides).  Returns -7 on a malformed template.      */
enum { SEG_LIT = -1 };
typedef struct { int group; int64_t lit_off, lit_len; } orc_seg;   /* literal bytes are tmpl[lit_off : lit_off+lit_len] or "$" (lit_off = -1) */

static int latin1_letter(unsigned c) {
  return (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == 0xAA || c == 0xB5 || c == 0xBA || (c >= 0xC0 && c <= 0xD6) ||
         (c >= 0xD8 && c <= 0xF6) || (c >= 0xF8 && c <= 0xFF);
}
static int t_name_start(unsigned c) { return c == '_' || latin1_letter(c); }
static int t_name_continue(unsigned c) { return c == '_' || latin1_letter(c) || (c >= '0' && c <= '9'); }

/* group index of a capture name (the switch of replace.go:436-453: named groups i >= 1), -1 when no case fires */
static int name_to_group(const orc_prog* P, const uint8_t* name, int64_t nl) {
  const uint32_t* w = P->w;
  const uint8_t* b = (const uint8_t*)(w + w[H_OFF_NAMES]);
  const size_t nb = (size_t)w[H_NAMES_WORDS] * 4;
  size_t pos = 0;
  const int groups = P->num_cap / 2;
  for (int g = 0; g < groups && pos < nb; g++) {
    const size_t len = b[pos++];
    if (g >= 1 && len > 0 && (int64_t)len == nl && memcmp(b + pos, name, len) == 0) return g;
    pos += len;
  }
  return -1;
}

static int template_parse(const orc_prog* P, const uint8_t* t, int64_t n, orc_seg* segs, int max_segs) {
  int ns = 0;
  const int groups = P->num_cap / 2;
#define ADD(g, o, l) do { if (ns >= max_segs) return -9; segs[ns].group = (g); segs[ns].lit_off = (o); segs[ns].lit_len = (l); ns++; } while (0)
#define ADD_GROUP(idx) do { if ((idx) >= 0 && (idx) < groups) ADD((int)(idx), 0, 0); } while (0)   /* CaptureByIndex: default -> nil */
  int64_t i = 0, literal_start = 0;
  while (i < n) {
    if (t[i] != '$') { i++; continue; }
    if (i > literal_start) ADD(SEG_LIT, literal_start, i - literal_start);
    if (i + 1 >= n) { ADD(SEG_LIT, -1, 1); i++; literal_start = i; continue; }
    const unsigned next = t[i + 1];
    if (next == '$') { ADD(SEG_LIT, -1, 1); i += 2; literal_start = i; }
    else if (next == '{') {
      int64_t close = -1;
      for (int64_t k = i; k < n; k++) if (t[k] == '}') { close = k; break; }
      if (close < 0) return -7;                                  /* unclosed ${ */
      const uint8_t* content = t + i + 2;
      const int64_t cl = close - (i + 2);
      if (cl == 0) return -7;                                    /* empty ${} */
      for (int64_t k = 0; k < cl; k++) if (content[k] >= 0x80) return -8;
      if (content[0] >= '0' && content[0] <= '9') {
        int64_t index = 0;
        for (int64_t k = 0; k < cl; k++) {
          if (content[k] < '0' || content[k] > '9') return -7;   /* mixed digits and non-digits */
          if (index < (1 << 24)) index = index * 10 + (content[k] - '0');
        }
        ADD_GROUP(index);
      } else {
        for (int64_t k = 0; k < cl; k++)
          if (k == 0 ? !t_name_start(content[k]) : !t_name_continue(content[k])) return -7;   /* invalid capture name */
        ADD_GROUP(name_to_group(P, content, cl));
      }
      i = close + 1; literal_start = i;
    } else if (next == '0') { ADD_GROUP(0); i += 2; literal_start = i; }
    else if (next >= '1' && next <= '9') {
      int64_t index = next - '0', consumed = 2;
      if (i + 2 < n && t[i + 2] >= '0' && t[i + 2] <= '9') { index = index * 10 + (t[i + 2] - '0'); consumed = 3; }
      ADD_GROUP(index);
      i += consumed; literal_start = i;
    } else if (t_name_start(next)) {
      int64_t end = i + 2;
      while (end < n && t_name_continue(t[end])) end++;
      ADD_GROUP(name_to_group(P, t + i + 1, end - (i + 1)));
      i = end; literal_start = i;
    } else { ADD(SEG_LIT, -1, 1); i++; literal_start = i; }
  }
  if (i > literal_start) ADD(SEG_LIT, literal_start, i - literal_start);
#undef ADD
#undef ADD_GROUP
  return ns;
}

/* One input.  Writes at most cap bytes; returns the length of the result, or < 0 (template). */
int64_t orc_replace_all(orc_prog* P, const uint8_t* in, int64_t l, const uint8_t* tmpl, int64_t tl, uint8_t* out, int64_t cap) {
  orc_seg segs[256];
  const int ns = template_parse(P, tmpl, tl, segs, 256);
  if (ns < 0) return ns;
  int64_t w = 0;
#define PUT(src, cnt) do { const int64_t c_ = (cnt); if (c_ > 0) { if (w + c_ <= cap) memcpy(out + w, (src), (size_t)c_); w += c_; } } while (0)
  int64_t offset = 0, last_end = 0;
  int64_t rec[MAX_TAGS];
  for (;;) {
    const uint8_t* rem = in + offset;
    const int64_t rl = l - offset;
    if (orc_find(P, rem, rl, rec) != 1) break;                              /* FindBytesReuse(remaining, &r) */
    const int64_t mlen = rec[1] - rec[0];
    int64_t midx;                                                           /* bytes.Index(remaining, match.Match) */
    if (mlen == 0) midx = 0;
    else {
      const uint8_t* q = (const uint8_t*)memmem(rem, (size_t)rl, rem + rec[0], (size_t)mlen);
      if (!q) break;
      midx = (int64_t)(q - rem);
    }
    const int64_t match_start = offset + midx, match_end = match_start + mlen;
    PUT(in + last_end, match_start - last_end);
    for (int s = 0; s < ns; s++) {
      if (segs[s].group == SEG_LIT) { if (segs[s].lit_off < 0) PUT("$", 1); else PUT(tmpl + segs[s].lit_off, segs[s].lit_len); }
      else { const int g = segs[s].group; if (rec[2 * g] >= 0) PUT(rem + rec[2 * g], rec[2 * g + 1] - rec[2 * g]); }
    }
    last_end = match_end;
    if (mlen > 0) offset = match_end;
    else if (match_end < l) offset = match_end + 1;
    else break;
  }
  PUT(in + last_end, l - last_end);
#undef PUT
  return w;
}

/* Batch: out_offs[n+1]; writes results back to back while they fit in cap; returns the total length or < 0. */
int64_t orc_replace_batch(orc_prog* P, const uint8_t* bytes, const uint64_t* offs, uint64_t n, const uint8_t* tmpl, int64_t tl,
                          uint8_t* out, int64_t cap, uint64_t* out_offs) {
  int64_t w = 0;
  for (uint64_t i = 0; i < n; i++) {
    out_offs[i] = (uint64_t)w;
    const int64_t r = orc_replace_all(P, bytes + offs[i], (int64_t)(offs[i + 1] - offs[i]), tmpl, tl, out + (w < cap ? w : cap), w < cap ? cap - w : 0);
    if (r < 0) return r;
    w += r;
  }
  out_offs[n] = (uint64_t)w;
  return w;
}
