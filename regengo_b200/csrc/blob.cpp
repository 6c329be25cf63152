// Serialise / parse the program blob (layout in blob.hpp).
#include "blob.hpp"

#include <cstring>

namespace rgx {

static void put64(std::vector<uint32_t>& w, uint64_t v) { w.push_back((uint32_t)v); w.push_back((uint32_t)(v >> 32)); }

std::vector<uint32_t> program_to_blob(const Program& P) {
  const Prog& prog = P.prog;
  const uint32_t n = (uint32_t)prog.inst.size();
  std::vector<uint32_t> w(H_WORDS_HDR, 0);
  w[H_MAGIC] = BLOB_MAGIC;
  w[H_VERSION] = BLOB_VERSION;
  w[H_NINST] = n;
  w[H_START] = (uint32_t)prog.start;
  w[H_NUMCAP] = (uint32_t)prog.num_cap;
  uint32_t f = 0;
  if (P.anchored) f |= F_ANCHORED;
  if (P.needs_backtracking) f |= F_NEEDS_BT;
  if (P.has_prefix) f |= F_HAS_PREFIX;
  if (P.match_memo) f |= F_MATCH_MEMO;
  if (P.find_memo) f |= F_FIND_MEMO;
  if (P.per_capture_ckpt) f |= F_PER_CAPTURE;
  if (P.has_captures) f |= F_HAS_CAPTURES;
  if (P.has_word_boundary) f |= F_WORD_BOUNDARY;
  w[H_FLAGS] = f;
  w[H_PREFIX] = P.prefix;
  w[H_MATCH_ENGINE] = (uint32_t)P.match_engine;
  w[H_FIND_ENGINE] = (uint32_t)P.find_engine;
  w[H_MINLEN] = (uint32_t)P.min_match_len;
  w[H_MAXLEN] = (uint32_t)P.max_match_len;
  w[H_LEFTOVER] = (uint32_t)P.default_max_leftover;
  w[H_MINBUF] = (uint32_t)P.min_buffer;

  // instructions
  w[H_OFF_INST] = (uint32_t)w.size();
  for (uint32_t i = 0; i < n; i++) {
    const Inst& in = prog.inst[i];
    uint32_t fl = 0;
    if (P.alt_ckpt[i]) fl |= IF_ALT_CKPT;
    if (P.greedy_loop[i]) fl |= IF_GREEDY_LOOP;
    if (P.unicode_class[i]) fl |= IF_UNICODE_CLASS;
    if (P.char_state[i]) fl |= IF_CHAR_STATE;
    w.push_back((uint32_t)in.op | (fl << 8));
    w.push_back(in.out);
    w.push_back(in.arg);
    w.push_back(in.op == InstRune1 && !in.rune.empty() ? (uint32_t)in.rune[0] : 0u);
  }
  // byte-class bitmaps
  w[H_OFF_CLASS] = (uint32_t)w.size();
  w.insert(w.end(), P.class_bits.begin(), P.class_bits.end());
  // rune ranges (kept for every Rune inst; only unicode classes consult them)
  w[H_OFF_RANGES] = (uint32_t)w.size();
  {
    std::vector<uint32_t> pairs;
    for (uint32_t i = 0; i < n; i++) {
      const Inst& in = prog.inst[i];
      uint32_t first = (uint32_t)(pairs.size() / 2), cnt = 0;
      if (in.op == InstRune) {
        if (in.rune.size() == 1) { pairs.push_back((uint32_t)in.rune[0]); pairs.push_back((uint32_t)in.rune[0]); cnt = 1; }
        else for (size_t k = 0; k + 1 < in.rune.size(); k += 2) { pairs.push_back((uint32_t)in.rune[k]); pairs.push_back((uint32_t)in.rune[k + 1]); cnt++; }
      }
      w.push_back(first);
      w.push_back(cnt);
    }
    w[H_NRANGE_PAIRS] = (uint32_t)(pairs.size() / 2);
    w.insert(w.end(), pairs.begin(), pairs.end());
  }
  // Thompson
  w[H_OFF_THOMPSON] = (uint32_t)w.size();
  put64(w, P.start_closure);
  put64(w, P.accept_mask);
  for (uint32_t i = 0; i < n; i++) put64(w, P.eps_after[i]);
  w.insert(w.end(), P.thompson_cond.begin(), P.thompson_cond.end());
  // TDFA
  const Tdfa& t = P.tdfa;
  if (t.built) {
    w[H_OFF_TDFA] = (uint32_t)w.size();
    size_t base = w.size();
    w.resize(base + T_WORDS_HDR, 0);
    const int ns = t.num_states;
    w[base + T_NSTATES] = (uint32_t)ns;
    w[base + T_NTAGS] = (uint32_t)t.num_tags;
    w[base + T_START_BEGIN] = (uint32_t)t.start_begin;
    w[base + T_START_ANY] = (uint32_t)t.start_any;
    w[base + T_NINIT_BEGIN] = (uint32_t)t.init_tags_begin.size();
    w[base + T_NINIT_ANY] = (uint32_t)t.init_tags_any.size();
    w[base + T_MAX_ACTS] = (uint32_t)t.max_actions;
    w[base + T_MAX_ACC_ACTS] = (uint32_t)t.max_accept_actions;
    for (int x : t.init_tags_begin) w.push_back((uint32_t)x);
    for (int x : t.init_tags_any) w.push_back((uint32_t)x);
    for (int32_t x : t.trans) w.push_back((uint32_t)x);
    for (uint8_t x : t.accept) w.push_back(x);
    for (uint8_t x : t.accept_eot) w.push_back(x);
    std::vector<uint32_t> off, acts;
    for (size_t i = 0; i < t.actions.size(); i++) {
      off.push_back((uint32_t)(acts.size() / 2));
      for (const TagAction& a : t.actions[i]) { acts.push_back((uint32_t)a.tag); acts.push_back((uint32_t)a.offset); }
    }
    off.push_back((uint32_t)(acts.size() / 2));
    w[base + T_NACTS] = (uint32_t)(acts.size() / 2);
    w.insert(w.end(), off.begin(), off.end());
    w.insert(w.end(), acts.begin(), acts.end());
    off.clear(); acts.clear();
    for (size_t i = 0; i < t.accept_actions.size(); i++) {
      off.push_back((uint32_t)(acts.size() / 2));
      for (const TagAction& a : t.accept_actions[i]) { acts.push_back((uint32_t)a.tag); acts.push_back((uint32_t)a.offset); }
    }
    off.push_back((uint32_t)(acts.size() / 2));
    w[base + T_NACC_ACTS] = (uint32_t)(acts.size() / 2);
    w.insert(w.end(), off.begin(), off.end());
    w.insert(w.end(), acts.begin(), acts.end());
  }
  // names
  {
    std::string bytes;
    for (const std::string& s : P.capture_names) { bytes.push_back((char)(s.size() > 255 ? 255 : s.size())); bytes.append(s.substr(0, 255)); }
    w[H_OFF_NAMES] = (uint32_t)w.size();
    size_t nw = (bytes.size() + 3) / 4;
    w[H_NAMES_WORDS] = (uint32_t)nw;
    size_t base = w.size();
    w.resize(base + nw, 0);
    if (!bytes.empty()) std::memcpy(&w[base], bytes.data(), bytes.size());
  }
  w[H_WORDS] = (uint32_t)w.size();
  return w;
}

#define NEED(cond) do { if (!(cond)) { err = "truncated or corrupt program blob (" #cond ")"; return false; } } while (0)

bool blob_to_program(const uint32_t* w, size_t nw, Program& P, std::string& err) {
  P = Program();
  NEED(nw >= H_WORDS_HDR);
  NEED(w[H_MAGIC] == BLOB_MAGIC);
  NEED(w[H_VERSION] == BLOB_VERSION);
  NEED(w[H_WORDS] == nw);
  const uint32_t n = w[H_NINST];
  NEED(n >= 1 && n < (1u << 20));
  P.prog.start = (int)w[H_START];
  P.prog.num_cap = (int)w[H_NUMCAP];
  NEED((uint32_t)P.prog.start < n && P.prog.num_cap >= 2 && P.prog.num_cap <= 1024);
  const uint32_t f = w[H_FLAGS];
  P.anchored = f & F_ANCHORED; P.needs_backtracking = f & F_NEEDS_BT; P.has_prefix = f & F_HAS_PREFIX;
  P.match_memo = f & F_MATCH_MEMO; P.find_memo = f & F_FIND_MEMO; P.per_capture_ckpt = f & F_PER_CAPTURE;
  P.has_captures = f & F_HAS_CAPTURES; P.has_word_boundary = f & F_WORD_BOUNDARY;
  P.prefix = (uint8_t)w[H_PREFIX];
  P.match_engine = (int)w[H_MATCH_ENGINE];
  P.find_engine = (int)w[H_FIND_ENGINE];
  NEED(P.match_engine == MATCH_BT || P.match_engine == MATCH_THOMPSON);
  NEED(P.find_engine >= FIND_NONE && P.find_engine <= FIND_TDFA);
  P.min_match_len = (int32_t)w[H_MINLEN]; P.max_match_len = (int32_t)w[H_MAXLEN];
  P.default_max_leftover = (int32_t)w[H_LEFTOVER]; P.min_buffer = (int32_t)w[H_MINBUF];

  size_t oi = w[H_OFF_INST], oc = w[H_OFF_CLASS], orr = w[H_OFF_RANGES], ot = w[H_OFF_THOMPSON];
  const uint32_t npairs = w[H_NRANGE_PAIRS];
  NEED(oi + (size_t)n * 4 <= nw && oc + (size_t)n * 8 <= nw && orr + (size_t)n * 2 + (size_t)npairs * 2 <= nw);
  NEED(ot + 4 + (size_t)n * 2 + (size_t)n * 8 <= nw);
  P.prog.inst.resize(n);
  P.alt_ckpt.assign(n, 0); P.greedy_loop.assign(n, 0); P.unicode_class.assign(n, 0); P.char_state.assign(n, 0);
  const uint32_t* pairs = w + orr + (size_t)n * 2;
  for (uint32_t i = 0; i < n; i++) {
    Inst& in = P.prog.inst[i];
    uint32_t w0 = w[oi + i * 4];
    in.op = (uint8_t)(w0 & 255);
    NEED(in.op <= InstRuneAnyNotNL);
    uint32_t fl = w0 >> 8;
    P.alt_ckpt[i] = fl & IF_ALT_CKPT ? 1 : 0; P.greedy_loop[i] = fl & IF_GREEDY_LOOP ? 1 : 0;
    P.unicode_class[i] = fl & IF_UNICODE_CLASS ? 1 : 0; P.char_state[i] = fl & IF_CHAR_STATE ? 1 : 0;
    in.out = w[oi + i * 4 + 1]; in.arg = w[oi + i * 4 + 2];
    NEED(in.out < n);
    if (in.op == InstAlt || in.op == InstAltMatch) NEED(in.arg < n);
    if (in.op == InstCapture) NEED(in.arg < (uint32_t)P.prog.num_cap);
    if (in.op == InstRune1) in.rune.push_back((int32_t)w[oi + i * 4 + 3]);
    if (in.op == InstRune) {
      uint32_t first = w[orr + i * 2], cnt = w[orr + i * 2 + 1];
      NEED((size_t)first + cnt <= npairs);
      for (uint32_t k = 0; k < cnt; k++) { in.rune.push_back((int32_t)pairs[(first + k) * 2]); in.rune.push_back((int32_t)pairs[(first + k) * 2 + 1]); }
    }
  }
  P.class_bits.assign(w + oc, w + oc + (size_t)n * 8);
  P.start_closure = (uint64_t)w[ot] | ((uint64_t)w[ot + 1] << 32);
  P.accept_mask = (uint64_t)w[ot + 2] | ((uint64_t)w[ot + 3] << 32);
  P.eps_after.resize(n);
  for (uint32_t i = 0; i < n; i++) P.eps_after[i] = (uint64_t)w[ot + 4 + i * 2] | ((uint64_t)w[ot + 5 + i * 2] << 32);
  P.thompson_cond.assign(w + ot + 4 + (size_t)n * 2, w + ot + 4 + (size_t)n * 2 + (size_t)n * 8);
  P.closures.assign(n, 0);

  if (w[H_OFF_TDFA]) {
    size_t b = w[H_OFF_TDFA];
    NEED(b + T_WORDS_HDR <= nw);
    Tdfa& t = P.tdfa;
    t.built = true;
    const uint32_t ns = w[b + T_NSTATES];
    NEED(ns >= 1 && ns <= 65535);
    t.num_states = (int)ns; t.num_tags = (int)w[b + T_NTAGS];
    NEED(t.num_tags >= 2 && t.num_tags <= 1024);
    t.start_begin = (int)w[b + T_START_BEGIN]; t.start_any = (int)w[b + T_START_ANY];
    NEED((uint32_t)t.start_begin < ns && (uint32_t)t.start_any < ns);
    t.max_actions = (int)w[b + T_MAX_ACTS]; t.max_accept_actions = (int)w[b + T_MAX_ACC_ACTS];
    const uint32_t nib = w[b + T_NINIT_BEGIN], nia = w[b + T_NINIT_ANY], nacts = w[b + T_NACTS], nacc = w[b + T_NACC_ACTS];
    size_t p = b + T_WORDS_HDR;
    size_t need = (size_t)nib + nia + (size_t)ns * 128 + ns * 2 + ((size_t)ns * 128 + 1) + (size_t)nacts * 2 + (ns + 1) + (size_t)nacc * 2;
    NEED(p + need <= nw);
    for (uint32_t i = 0; i < nib; i++) { NEED(w[p] < (uint32_t)t.num_tags); t.init_tags_begin.push_back((int)w[p++]); }
    for (uint32_t i = 0; i < nia; i++) { NEED(w[p] < (uint32_t)t.num_tags); t.init_tags_any.push_back((int)w[p++]); }
    t.trans.resize((size_t)ns * 128);
    for (size_t i = 0; i < (size_t)ns * 128; i++) { int32_t v = (int32_t)w[p++]; NEED(v >= -1 && v < (int32_t)ns); t.trans[i] = v; }
    t.accept.resize(ns); for (uint32_t i = 0; i < ns; i++) t.accept[i] = (uint8_t)w[p++];
    t.accept_eot.resize(ns); for (uint32_t i = 0; i < ns; i++) t.accept_eot[i] = (uint8_t)w[p++];
    const uint32_t* off = w + p; p += (size_t)ns * 128 + 1;
    const uint32_t* acts = w + p; p += (size_t)nacts * 2;
    t.actions.resize((size_t)ns * 128);
    for (size_t i = 0; i < (size_t)ns * 128; i++) {
      NEED(off[i] <= off[i + 1] && off[i + 1] <= nacts);
      for (uint32_t k = off[i]; k < off[i + 1]; k++) { NEED(acts[k * 2] < (uint32_t)t.num_tags); t.actions[i].push_back({(int)acts[k * 2], (int)acts[k * 2 + 1]}); }
    }
    const uint32_t* aoff = w + p; p += ns + 1;
    const uint32_t* aacts = w + p; p += (size_t)nacc * 2;
    t.accept_actions.resize(ns);
    for (uint32_t i = 0; i < ns; i++) {
      NEED(aoff[i] <= aoff[i + 1] && aoff[i + 1] <= nacc);
      for (uint32_t k = aoff[i]; k < aoff[i + 1]; k++) { NEED(aacts[k * 2] < (uint32_t)t.num_tags); t.accept_actions[i].push_back({(int)aacts[k * 2], (int)aacts[k * 2 + 1]}); }
    }
  }
  if (P.find_engine == FIND_TDFA) NEED(P.tdfa.built);
  {
    size_t on = w[H_OFF_NAMES], nn = w[H_NAMES_WORDS];
    NEED(on + nn <= nw);
    const uint8_t* b = reinterpret_cast<const uint8_t*>(w + on);
    size_t nb = nn * 4, pos = 0;
    int groups = P.prog.num_cap / 2;
    for (int g = 0; g < groups && pos < nb && P.has_captures; g++) {
      size_t len = b[pos++];
      NEED(pos + len <= nb);
      P.capture_names.emplace_back(reinterpret_cast<const char*>(b + pos), len);
      pos += len;
    }
  }
  return true;
}

}  // namespace rgx
