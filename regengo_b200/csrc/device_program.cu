// Host-side packing of a compiled Program into the device image (layout: device_program.cuh).
#include "device_program.cuh"

#include <cstring>
#include <map>

namespace rgx {

namespace {

struct Bits256 {
  uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  void set(uint32_t c) { w[(c & 255) >> 5] |= 1u << (c & 31); }
  bool all() const { for (int i = 0; i < 8; i++) if (w[i] != 0xFFFFFFFFu) return false; return true; }
};

// Bytes that can be the first byte consumed by an attempt of the goto-machine, and whether the
// attempt can reach Match without consuming anything.  Empty-width assertions are assumed to pass
// (over-approximation: the set is only used to skip starts that cannot match).
void first_bytes_bt(const Program& P, Bits256& first, bool& nullable) {
  const Prog& prog = P.prog;
  std::vector<uint8_t> seen(prog.inst.size(), 0);
  std::vector<int> st{prog.start};
  nullable = false;
  while (!st.empty()) {
    int pc = st.back(); st.pop_back();
    if (seen[pc]) continue;
    seen[pc] = 1;
    const Inst& in = prog.inst[pc];
    switch (in.op) {
      case InstAlt: case InstAltMatch: st.push_back((int)in.out); st.push_back((int)in.arg); break;
      case InstNop: case InstCapture: case InstEmptyWidth: st.push_back((int)in.out); break;
      case InstMatch: nullable = true; break;
      case InstFail: break;
      case InstRune1: {
        int32_t r = in.rune.empty() ? 0 : in.rune[0];
        if (r < 128) first.set((uint32_t)r);
        else if (r < 0x800) first.set(0xC0u | ((uint32_t)r >> 6));
        else if (r < 0x10000) first.set(0xE0u | ((uint32_t)r >> 12));
        else first.set(0xF0u | ((uint32_t)r >> 18));
        break;
      }
      case InstRune:
        for (int k = 0; k < 8; k++) first.w[k] |= P.class_bits[(size_t)pc * 8 + k];
        if (P.unicode_class[pc]) for (uint32_t c = 128; c < 256; c++) first.set(c);
        break;
      case InstRuneAny: for (uint32_t c = 0; c < 256; c++) first.set(c); break;
      case InstRuneAnyNotNL: for (uint32_t c = 0; c < 256; c++) if (c != '\n') first.set(c); break;
    }
  }
}

}  // namespace

void pack_program(const Program& P, std::vector<uint32_t>& w, DevMeta& m) {
  std::memset(&m, 0, sizeof(m));
  const Prog& prog = P.prog;
  const uint32_t n = (uint32_t)prog.inst.size();
  m.n_inst = (int32_t)n;
  m.start = prog.start;
  m.num_cap = prog.num_cap;
  uint32_t f = 0;
  if (P.anchored) f |= F_ANCHORED;
  if (P.needs_backtracking) f |= F_NEEDS_BT;
  if (P.has_prefix) f |= F_HAS_PREFIX;
  if (P.match_memo) f |= F_MATCH_MEMO;
  if (P.find_memo) f |= F_FIND_MEMO;
  if (P.per_capture_ckpt) f |= F_PER_CAPTURE;
  if (P.has_captures) f |= F_HAS_CAPTURES;
  m.flags = (int32_t)f;
  m.prefix = P.prefix;
  m.match_engine = P.match_engine;
  m.find_engine = P.find_engine;
  w.clear();
  auto align4 = [&]() { while (w.size() % 4) w.push_back(0); };

  m.off_inst = (uint32_t)w.size();
  for (uint32_t i = 0; i < n; i++) {
    const Inst& in = prog.inst[i];
    uint32_t fl = 0;
    if (P.alt_ckpt[i]) fl |= IF_ALT_CKPT;
    if (P.greedy_loop[i]) fl |= IF_GREEDY_LOOP;
    if (P.unicode_class[i]) fl |= IF_UNICODE_CLASS;
    if (P.char_state[i]) fl |= IF_CHAR_STATE;
    if (in.op == InstAlt) m.n_alt++;
    if (in.op == InstCapture) m.n_capinst++;
    if (in.op == InstEmptyWidth) m.n_empty++;
    w.push_back((uint32_t)in.op | (fl << 8));
    w.push_back(in.out);
    w.push_back(in.arg);
    w.push_back(in.op == InstRune1 && !in.rune.empty() ? (uint32_t)in.rune[0] : 0u);
  }
  m.off_cls = (uint32_t)w.size();
  w.insert(w.end(), P.class_bits.begin(), P.class_bits.end());
  m.off_th_eps = (uint32_t)w.size();
  uint64_t char_mask = 0;
  for (uint32_t i = 0; i < n; i++) {
    w.push_back((uint32_t)P.eps_after[i]);
    w.push_back((uint32_t)(P.eps_after[i] >> 32));
    if (P.char_state[i] && i < 64) char_mask |= 1ull << i;
  }
  m.off_th_cond = (uint32_t)w.size();
  w.insert(w.end(), P.thompson_cond.begin(), P.thompson_cond.end());
  m.th_start_lo = (uint32_t)P.start_closure; m.th_start_hi = (uint32_t)(P.start_closure >> 32);
  m.th_accept_lo = (uint32_t)P.accept_mask; m.th_accept_hi = (uint32_t)(P.accept_mask >> 32);
  m.th_char_lo = (uint32_t)char_mask; m.th_char_hi = (uint32_t)(char_mask >> 32);
  // unicode ranges
  m.off_rng_idx = (uint32_t)w.size();
  {
    std::vector<uint32_t> pairs;
    for (uint32_t i = 0; i < n; i++) {
      const Inst& in = prog.inst[i];
      uint32_t first = (uint32_t)(pairs.size() / 2), cnt = 0;
      if (in.op == InstRune && P.unicode_class[i])
        for (size_t k = 0; k + 1 < in.rune.size(); k += 2) { pairs.push_back((uint32_t)in.rune[k]); pairs.push_back((uint32_t)in.rune[k + 1]); cnt++; }
      w.push_back(first);
      w.push_back(cnt);
    }
    m.off_rng_pairs = (uint32_t)w.size();
    w.insert(w.end(), pairs.begin(), pairs.end());
  }
  align4();
  m.match_words = (uint32_t)w.size();   // everything MatchBytes reads lies before this point

  // TDFA tables
  Bits256 first;
  bool nullable = false;
  const Tdfa& t = P.tdfa;
  if (t.built && P.find_engine == FIND_TDFA) {
    m.t_ns = t.num_states; m.t_ntags = t.num_tags;
    m.t_start_begin = t.start_begin; m.t_start_any = t.start_any;
    m.t_n_init_begin = (int32_t)t.init_tags_begin.size(); m.t_n_init_any = (int32_t)t.init_tags_any.size();
    // de-duplicated action lists; list 0 = empty
    std::map<std::vector<uint32_t>, uint32_t> ids;
    std::vector<std::vector<uint32_t>> lists;
    auto list_id = [&](const std::vector<TagAction>& a) -> uint32_t {
      std::vector<uint32_t> key;
      for (const TagAction& x : a) key.push_back((uint32_t)x.tag | ((uint32_t)x.offset << 16));
      auto it = ids.find(key);
      if (it != ids.end()) return it->second;
      uint32_t id = (uint32_t)lists.size();
      ids[key] = id; lists.push_back(key);
      return id;
    };
    list_id({});
    m.off_t_trans = (uint32_t)w.size();
    for (size_t i = 0; i < (size_t)t.num_states * 128; i++) {
      uint32_t nx = t.trans[i] < 0 ? TDFA_NONE : (uint32_t)t.trans[i];
      uint32_t al = t.trans[i] < 0 ? 0 : list_id(t.actions[i]);
      w.push_back(nx | (al << 16));
    }
    m.off_t_accept = (uint32_t)w.size();
    for (int s = 0; s < t.num_states; s++) {
      uint32_t fl = (t.accept[s] ? 1u : 0u) | (t.accept_eot[s] ? 2u : 0u);
      w.push_back(fl | (list_id(t.accept_actions[s]) << 16));
    }
    m.off_t_alist_off = (uint32_t)w.size();
    {
      uint32_t pos = 0;
      for (auto& l : lists) { w.push_back(pos); pos += (uint32_t)l.size(); }
      w.push_back(pos);
    }
    m.off_t_alist = (uint32_t)w.size();
    for (auto& l : lists) w.insert(w.end(), l.begin(), l.end());
    m.off_t_init = (uint32_t)w.size();
    for (int x : t.init_tags_begin) w.push_back((uint32_t)x);
    for (int x : t.init_tags_any) w.push_back((uint32_t)x);
    align4();
    // start filter from the tables: bytes with a transition out of startStateAny
    nullable = t.accept[t.start_any] || t.accept_eot[t.start_any];
    for (uint32_t c = 0; c < 128; c++) if (t.trans[(size_t)t.start_any * 128 + c] >= 0) first.set(c);
    int s = t.start_any;
    std::vector<int> prefix_states{s};   // prefix_states[i] = state after i prefix bytes
    std::vector<uint32_t> prefix_lists;  // transition list fired by prefix byte i
    while (m.prefix_len < MAX_PREFIX && !t.accept[s] && !t.accept_eot[s]) {
      int only = -1, cnt = 0;
      for (int c = 0; c < 128; c++) if (t.trans[(size_t)s * 128 + c] >= 0) { only = c; cnt++; }
      if (cnt != 1) break;
      m.prefix_bytes[m.prefix_len++] = (uint8_t)only;
      prefix_lists.push_back(list_id(t.actions[(size_t)s * 128 + only]));
      s = t.trans[(size_t)s * 128 + only];
      prefix_states.push_back(s);
    }
    // scan6 walk image (layout: device_program.cuh).  The walk distinguishes only two kinds of step: CHEAP ones,
    // whose whole effect is "go to that row", and EVENTS, which it logs and interprets when the walk is over.
    // start filter of scan6: a literal first byte (prefix), or a first-byte SET of at most two ASCII ranges (\d, [a-z] ...)
    int n_rng = 0;
    uint8_t rlo[2] = {0, 0}, rhi[2] = {0, 0};
    bool set_ok = m.prefix_len == 0;
    for (uint32_t c = 0; set_ok && c < 128; c++) {
      const bool in = (first.w[c >> 5] >> (c & 31)) & 1u;
      const bool prev = c > 0 && ((first.w[(c - 1) >> 5] >> ((c - 1) & 31)) & 1u);
      if (in && !prev) { if (n_rng == 2) { set_ok = false; break; } rlo[n_rng] = (uint8_t)c; rhi[n_rng] = (uint8_t)c; n_rng++; }
      else if (in) rhi[n_rng - 1] = (uint8_t)c;
    }
    set_ok = set_ok && n_rng >= 1;
    if ((m.prefix_len >= 1 || set_ok) && !nullable && t.start_begin == t.start_any && t.init_tags_begin == t.init_tags_any &&
        lists.size() <= 1023 && t.num_tags <= 16) {
      std::vector<uint32_t> rows((size_t)t.num_states * 256, S6_DEAD);
      std::map<std::pair<int, uint32_t>, uint32_t> desc_ids;
      std::vector<uint32_t> desc{0u};                 // descriptor 0: "no event" (cheap cells carry index 0)
      std::vector<uint32_t> fent{0u, 0u, 0u, 0u};     // S6_FENT entry words per descriptor
      bool fits = true;
      for (int st = 0; st < t.num_states; st++) {
        const bool acc_s = t.accept[st] || t.accept_eot[st];
        for (int c = 0; c < 128; c++) {
          const size_t i = (size_t)st * 128 + c;
          if (t.trans[i] < 0) continue;
          const int nx = t.trans[i];
          const uint32_t tl = list_id(t.actions[i]);
          const bool acc_n = t.accept[nx] || t.accept_eot[nx];
          if (tl == 0 && (nx == st || (!acc_s && !acc_n))) { rows[(size_t)st * 256 + c] = (uint32_t)nx * 1024u; continue; }
          auto key = std::make_pair(nx, tl);
          auto it = desc_ids.find(key);
          uint32_t id;
          if (it != desc_ids.end()) id = it->second;
          else {
            id = (uint32_t)desc.size();
            desc_ids[key] = id;
            // the event's effect on the tags, in the order the reference applies it: the transition's actions at the
            // step's position, then (while the next state accepts) that state's accept actions at the end of its run
            std::vector<uint32_t> ent;
            for (const TagAction& x : t.actions[i]) {
              if (x.tag >= 16 || x.offset > 255 || x.offset < 0) fits = false;
              ent.push_back((uint32_t)x.tag * 128u | ((uint32_t)x.offset << 16));
            }
            if (acc_n)
              for (const TagAction& x : t.accept_actions[nx]) {
                if (x.tag >= 16 || x.offset > 255 || x.offset < 0) fits = false;
                ent.push_back((uint32_t)x.tag * 128u | ((uint32_t)x.offset << 16) | S6_ENT_ACCEPT);
              }
            if (ent.size() > S6_FENT) fits = false;
            desc.push_back((t.accept[nx] ? S6_ACC : 0u) | (t.accept_eot[nx] ? S6_ACC_EOT : 0u) | ((uint32_t)std::min<size_t>(ent.size(), S6_FENT) << 24));
            ent.resize(S6_FENT, 0u);
            fent.insert(fent.end(), ent.begin(), ent.end());
          }
          rows[(size_t)st * 256 + c] = (id << 22) | ((uint32_t)nx * 1024u);
        }
      }
      const size_t ndesc = desc.size();
      const size_t bytes = (rows.size() + desc.size() * 8 + 64) * 4;
      if (fits && ndesc <= 1022 && (size_t)t.num_states * 1024 < (1u << 22) && bytes <= S6_IMAGE_LIMIT) {
        align4();
        m.w6_off = (uint32_t)w.size();
        w.insert(w.end(), rows.begin(), rows.end());
        // descriptor table: 8 words per descriptor = {flags | n << 24, S6_FENT entries, padding}: one shift to address
        align4();
        m.w6_desc = (uint32_t)w.size() - m.w6_off;
        for (size_t d = 0; d < ndesc; d++) {
          w.push_back(desc[d]);
          for (uint32_t q = 0; q < S6_FENT; q++) w.push_back(fent[d * S6_FENT + q]);
          for (uint32_t q = 1 + S6_FENT; q < 8; q++) w.push_back(0);
        }
        m.w6_ndesc = (int32_t)ndesc;
        m.w6_fent = m.w6_desc + 1;
        m.w6_init = (uint32_t)w.size() - m.w6_off;
        for (int x : t.init_tags_any) w.push_back((uint32_t)x);
        align4();
        m.w6_words = (uint32_t)w.size() - m.w6_off;
        m.w6_ok = 1;
        // the most events one walk can log: longest path from the start state counting event edges (a cycle through
        // an event edge: no bound).  The scan picks its per-candidate log size from it.
        {
          std::vector<int> dist((size_t)t.num_states, -1);
          dist[t.start_any] = 0;
          bool changed = true;
          int rounds = 0;
          while (changed && rounds <= t.num_states + 1) {
            changed = false;
            rounds++;
            for (int st = 0; st < t.num_states; st++) {
              if (dist[st] < 0) continue;
              for (int c = 0; c < 128; c++) {
                const uint32_t cell = rows[(size_t)st * 256 + c];
                if (cell >= S6_DEAD) continue;
                const int nx = (int)((cell & 0x3FFFFFu) / 1024u), wgt = cell >= S6_EVMIN ? 1 : 0;
                if (dist[st] + wgt > dist[nx]) { dist[nx] = dist[st] + wgt; changed = true; }
              }
            }
          }
          int mx = 0;
          for (int d : dist) mx = std::max(mx, d);
          m.w6_maxev = changed ? 255 : std::min(mx, 255);
        }
        if (m.prefix_len == 0) { m.w6_nrng = n_rng; for (int q = 0; q < 2; q++) { m.w6_rlo[q] = rlo[q]; m.w6_rhi[q] = rhi[q]; } }
      }
    }
  } else {
    first_bytes_bt(P, first, nullable);
    int pc = prog.start;
    for (size_t guard = 0; guard <= n && m.prefix_len < MAX_PREFIX; guard++) {
      const Inst& in = prog.inst[pc];
      if (in.op == InstNop || in.op == InstCapture) { pc = (int)in.out; continue; }
      if (in.op == InstRune1 && in.rune.size() == 1 && in.rune[0] < 128) { m.prefix_bytes[m.prefix_len++] = (uint8_t)in.rune[0]; pc = (int)in.out; continue; }
      break;
    }
  }
  m.nullable = nullable ? 1 : 0;
  m.off_first = (uint32_t)w.size();
  for (int k = 0; k < 8; k++) w.push_back(first.w[k]);
  align4();
  if (nullable) { m.gen_kind = GEN_ALL; m.prefix_len = 0; }
  else if (m.prefix_len > 0) m.gen_kind = GEN_PREFIX;
  else if (!first.all()) m.gen_kind = GEN_BYTESET;
  else m.gen_kind = GEN_ALL;
  // "Atomic" greedy loops: an Alt A whose body is one ASCII class instruction C looping straight back to
  // A, and whose exit (reached only from A) leads through captures/nops to a literal byte b outside C.
  // Of the alternatives the reference stacks for such a loop only the NEWEST is the loop's exit; the older
  // ones (shorter run lengths) resume at a byte of C and die at once on b.  A FindAll attempt -- where only
  // success, end and captures of the successful path are observable -- therefore keeps one stack entry
  // per loop and overwrites it each iteration (engines.cuh, MODE_FINDALL without memoisation).
  std::vector<int> indeg_all(n, 0);
  for (uint32_t i = 0; i < n; i++) {
    const Inst& in = prog.inst[i];
    if (in.op == InstFail || in.op == InstMatch) continue;
    indeg_all[in.out]++;
    if (in.op == InstAlt || in.op == InstAltMatch) indeg_all[in.arg]++;
  }
  for (uint32_t i = 0; i < n; i++) {
    const Inst& a = prog.inst[i];
    if (a.op != InstAlt || indeg_all[a.arg] != 1) continue;
    const uint32_t L = a.out;
    if (L >= n || prog.inst[L].op != InstRune || P.unicode_class[L] || prog.inst[L].out != i) continue;
    uint32_t pc = a.arg;
    size_t guard = 0;
    while (guard++ <= n && (prog.inst[pc].op == InstCapture || prog.inst[pc].op == InstNop)) pc = prog.inst[pc].out;
    const Inst& lit = prog.inst[pc];
    if (lit.op != InstRune1 || lit.rune.size() != 1 || lit.rune[0] >= 128) continue;
    const uint32_t b = (uint32_t)lit.rune[0];
    if ((P.class_bits[(size_t)L * 8 + (b >> 5)] >> (b & 31)) & 1u) continue;
    w[m.off_inst + 4 * i] |= (uint32_t)IF_ATOMIC_LOOP << 8;
  }
  // Straight-line program, or at least a straight-line PREFIX: captures and nops between single-byte ASCII class steps
  // from the start instruction on -- every match has to pass them in this order.  sl_n > 0: the whole program is such a
  // line (then Match); slp_n: the steps before the first instruction that is not (an Alt, an EmptyWidth, Match, ...).
  if (P.find_engine == FIND_BT) {
    std::vector<Bits256> classes;
    std::vector<uint8_t> steps;
    std::vector<int> cap_at(MAX_CAPS, -1);   // capture slot -> bytes consumed before its Capture instruction
    int pc = prog.start;
    bool ok = true, done = false;
    size_t guard = 0;
    while (ok && !done && guard++ <= n) {
      const Inst& in = prog.inst[pc];
      Bits256 set;
      bool consumes = false;
      switch (in.op) {
        case InstCapture:
          if (in.arg < (uint32_t)MAX_CAPS) cap_at[in.arg] = (int)steps.size();
          pc = (int)in.out; break;
        case InstNop: pc = (int)in.out; break;
        case InstRune1:
          if (in.rune.size() != 1 || in.rune[0] >= 128) { ok = false; break; }
          set.set((uint32_t)in.rune[0]); consumes = true; break;
        case InstRune:
          if (P.unicode_class[pc]) { ok = false; break; }
          for (int k = 0; k < 8; k++) set.w[k] = P.class_bits[(size_t)pc * 8 + k];
          if (set.w[4] | set.w[5] | set.w[6] | set.w[7]) { ok = false; break; }   // ASCII sets only
          consumes = true; break;
        case InstMatch: done = true; break;
        default: ok = false; break;
      }
      if (ok && consumes) {
        size_t k = 0;
        for (; k < classes.size(); k++) if (std::memcmp(classes[k].w, set.w, sizeof(set.w)) == 0) break;
        if (k == classes.size() && classes.size() == 8) { ok = false; break; }
        if (steps.size() >= 32) { ok = false; break; }
        if (k == classes.size()) classes.push_back(set);
        steps.push_back((uint8_t)k);
        pc = (int)in.out;
      }
    }
    const bool whole = ok && done && !steps.empty() && m.n_alt == 0 && m.n_empty == 0;
    if (whole || steps.size() >= 2) {
      m.slp_n = (int32_t)steps.size();
      m.sl_n = whole ? (int32_t)steps.size() : 0;
      m.sl_ncls = (int32_t)classes.size();
      for (size_t i = 0; i < steps.size(); i++) m.sl_cls[i] = steps[i];
      if (whole) {
        // every capture of a straight-line program sits at a fixed distance from the match start (slots 0 / 1: the match)
        m.sl_caps_ok = prog.num_cap <= MAX_CAPS ? 1 : 0;
        for (int slot = 0; slot < prog.num_cap && slot < MAX_CAPS; slot++) {
          const int at = slot == 0 ? 0 : slot == 1 ? (int)steps.size() : cap_at[slot];
          if (at < 0) m.sl_caps_ok = 0; else m.sl_cap[slot] = (uint8_t)at;
        }
      }
      m.off_sl_cm = (uint32_t)w.size();
      for (uint32_t c0 = 0; c0 < 256; c0 += 4) {
        uint32_t word = 0;
        for (uint32_t j = 0; j < 4; j++) {
          const uint32_t c = c0 + j;
          uint32_t bits = 0;
          for (size_t k = 0; k < classes.size(); k++) if ((classes[k].w[c >> 5] >> (c & 31)) & 1u) bits |= 1u << k;
          word |= bits << (8 * j);
        }
        w.push_back(word);
      }
      align4();
    }
  }
  m.image_words = (uint32_t)w.size();

  // Recognise  (cap|nop)* C+ (cap|nop)* b ...  (greedy loop over an ASCII class C, then a literal byte
  // b outside C).  An attempt at s then succeeds iff the maximal C-run from s is followed by b and the
  // rest matches after it -- the same outcome for every s of the run (kernels_btrun.cuh has the proof).
  if (P.find_engine == FIND_BT && !P.anchored && !P.find_memo) {
    std::vector<int> indeg(n, 0);
    for (uint32_t i = 0; i < n; i++) {
      const Inst& in = prog.inst[i];
      if (in.op == InstFail || in.op == InstMatch) continue;
      indeg[in.out]++;
      if (in.op == InstAlt || in.op == InstAltMatch) indeg[in.arg]++;
    }
    int pc = prog.start;
    uint32_t start_caps = 1u;  // captures[0] = searchStart
    bool ok = true;
    size_t guard = 0;
    while (ok && guard++ <= n && (prog.inst[pc].op == InstCapture || prog.inst[pc].op == InstNop)) {
      if (prog.inst[pc].op == InstCapture) { if (prog.inst[pc].arg < 32) start_caps |= 1u << prog.inst[pc].arg; else ok = false; }
      if (pc != prog.start && indeg[pc] != 1) ok = false;
      pc = (int)prog.inst[pc].out;
    }
    const int L = pc;
    ok = ok && prog.inst[L].op == InstRune && !P.unicode_class[L] && indeg[L] == (L == prog.start ? 1 : 2);
    int A = ok ? (int)prog.inst[L].out : 0;
    ok = ok && prog.inst[A].op == InstAlt && (int)prog.inst[A].out == L && indeg[A] == 1 && A != L;
    if (ok) {
      pc = (int)prog.inst[A].arg;
      guard = 0;
      while (ok && guard++ <= n && (prog.inst[pc].op == InstCapture || prog.inst[pc].op == InstNop)) {
        if (indeg[pc] != 1) ok = false;
        pc = (int)prog.inst[pc].out;
      }
      const Inst& lit = prog.inst[pc];
      ok = ok && lit.op == InstRune1 && lit.rune.size() == 1 && lit.rune[0] < 128 && indeg[pc] == 1;
      if (ok) {
        const uint32_t b = (uint32_t)lit.rune[0];
        if ((P.class_bits[(size_t)L * 8 + (b >> 5)] >> (b & 31)) & 1u) ok = false;   // b must not be in C
        if (ok) { m.run_ok = 1; m.run_class_pc = L; m.run_lit = (int32_t)b; m.run_start_caps = start_caps; m.run_resume_pc = (int)prog.inst[A].arg; }
      }
    }
    // Straight-line continuation after the leading loop?  Elements: Capture, ASCII literal byte, ASCII class,
    // and "C+" = class instruction L followed by Alt(out = L, arg = exit) with L looping straight back to the
    // Alt, where the loop is ATOMIC: the exit leads through captures/nops either to a literal byte outside C
    // or to Match.  Then a shorter run of the loop can never help (it would resume on a byte of C where a byte
    // outside C is required; before Match the greedy choice succeeds outright), so the goto-machine's first
    // successful path is the greedy one and every other path fails: direct evaluation gives the same verdict,
    // the same end and the same captures (instructions.go:331-457, find.go:351-466).
    if (m.run_ok) {
      std::vector<uint32_t> lin;
      int pc = m.run_resume_pc;
      bool ok = true, done = false;
      std::vector<uint8_t> seen(n, 0);
      while (ok && !done && lin.size() < 39) {
        if (seen[pc]) { ok = false; break; }
        seen[pc] = 1;
        const Inst& in = prog.inst[pc];
        switch (in.op) {
          case InstNop: pc = (int)in.out; break;
          case InstCapture: if (in.arg >= 32) ok = false; lin.push_back(LIN_CAP | (in.arg << 8)); pc = (int)in.out; break;
          case InstRune1:
            if (in.rune.size() != 1 || in.rune[0] >= 128) { ok = false; break; }
            lin.push_back(LIN_LIT | ((uint32_t)in.rune[0] << 8)); pc = (int)in.out; break;
          case InstRune: {
            if (P.unicode_class[pc]) { ok = false; break; }
            lin.push_back(LIN_CLS | ((uint32_t)pc << 8));
            const int A2 = (int)in.out;
            const Inst& a2 = prog.inst[A2];
            if (a2.op == InstAlt && (int)a2.out == pc && indeg[A2] == 1 && indeg[pc] == 2) {
              // C+ : is the loop atomic?
              int q = (int)a2.arg;
              size_t g2 = 0;
              while (g2++ <= n && (prog.inst[q].op == InstCapture || prog.inst[q].op == InstNop)) q = (int)prog.inst[q].out;
              const Inst& nx = prog.inst[q];
              bool atomic = nx.op == InstMatch;
              if (nx.op == InstRune1 && nx.rune.size() == 1 && nx.rune[0] < 128) {
                const uint32_t b2 = (uint32_t)nx.rune[0];
                atomic = !((P.class_bits[(size_t)pc * 8 + (b2 >> 5)] >> (b2 & 31)) & 1u);
              }
              if (!atomic) { ok = false; break; }
              lin.push_back(LIN_LOOP | ((uint32_t)pc << 8));
              seen[A2] = 1;
              pc = (int)a2.arg;
            } else {
              pc = A2;
            }
            break;
          }
          case InstMatch: lin.push_back(LIN_MATCH); done = true; break;
          default: ok = false; break;
        }
      }
      if (ok && done) {
        m.lin_n = (int32_t)lin.size();
        for (size_t i = 0; i < lin.size(); i++) m.lin[i] = lin[i];
      }
    }
  }
}

}  // namespace rgx
