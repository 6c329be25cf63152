// See program.hpp.  Every function cites the reference lines it restates.
#include "program.hpp"

#include <algorithm>
#include <deque>
#include <set>
#include <sstream>

namespace rgx {

// ---------------------------------------------------------------------------------------
// analysis.go
static void walk_capture_names(const Regexp* re, std::map<int, std::string>& cap_map, int& max_cap) {
  // analysis.go:36-66 extractCaptureNames (first occurrence of each Cap wins)
  if (re->op == OpCapture) {
    if (!cap_map.count(re->cap)) {
      cap_map[re->cap] = re->name;
      if (re->cap > max_cap) max_cap = re->cap;
    }
  }
  for (const Regexp* s : re->sub) walk_capture_names(s, cap_map, max_cap);
}

static bool needs_backtracking(const Prog& p) {  // analysis.go:78-90
  for (const Inst& i : p.inst) if (i.op == InstAlt) return true;
  return false;
}

static bool is_anchored(const Prog& p) {  // analysis.go:117-124
  if (p.inst.empty()) return false;
  const Inst& s = p.inst[p.start];
  return s.op == InstEmptyWidth && (s.arg & EmptyBeginText) != 0;
}

static bool has_word_boundary(const Prog& p) {  // analysis.go:127-142
  for (const Inst& i : p.inst)
    if (i.op == InstEmptyWidth && (i.arg & (EmptyWordBoundary | EmptyNoWordBoundary))) return true;
  return false;
}

static bool has_unicode_char_class(const Prog& p) {  // analysis.go:146-163
  for (const Inst& i : p.inst)
    if (i.op == InstRune || i.op == InstRune1)
      for (int32_t r : i.rune) if (r >= 128) return true;
  return false;
}

static bool is_simple_loop(const Prog& p, int start_idx) {  // analysis.go:212-246
  const Inst& inst = p.inst[start_idx];
  std::deque<int> queue = {(int)inst.out, (int)inst.arg};
  std::set<int> visited = {start_idx};
  while (!queue.empty()) {
    int curr = queue.front(); queue.pop_front();
    if (curr == start_idx) return true;
    if (visited.count(curr)) continue;
    visited.insert(curr);
    const Inst& ci = p.inst[curr];
    if (ci.op == InstAlt) return false;
    if (ci.op != InstMatch && ci.op != InstFail) queue.push_back((int)ci.out);
  }
  return false;
}

static bool reaches(const Prog& p, int start, int target) {  // analysis.go:249-278
  std::deque<int> queue = {start};
  std::set<int> visited = {start};
  while (!queue.empty()) {
    int curr = queue.front(); queue.pop_front();
    if (curr == target) return true;
    const Inst& inst = p.inst[curr];
    std::vector<int> next;
    if (inst.op == InstAlt) next = {(int)inst.out, (int)inst.arg};
    else if (inst.op != InstMatch && inst.op != InstFail) next = {(int)inst.out};
    for (int n : next)
      if (!visited.count(n)) { visited.insert(n); queue.push_back(n); }
  }
  return false;
}

static bool detect_complexity(const Prog& p) {  // analysis.go:168-209
  std::vector<int> alts;
  for (size_t i = 0; i < p.inst.size(); i++) if (p.inst[i].op == InstAlt) alts.push_back((int)i);
  if (alts.size() < 2) return false;
  std::vector<int> simple;
  for (int a : alts) if (is_simple_loop(p, a)) simple.push_back(a);
  for (int head : simple)
    for (int other : alts) {
      if (head == other) continue;
      if (reaches(p, head, other) && reaches(p, other, head)) return true;
    }
  return false;
}

static bool has_end_anchor(const Prog& p) {  // analysis.go:317-331
  for (const Inst& i : p.inst) if (i.op == InstEmptyWidth && (i.arg & EmptyEndText)) return true;
  return false;
}

static bool nested_quantifiers(const Regexp* re, int depth) {  // analysis.go:341-369
  if (re == nullptr) return false;
  bool is_q = false;
  switch (re->op) {
    case OpStar: case OpPlus: case OpQuest: case OpRepeat:
      is_q = true;
      if (depth > 0) return true;
  }
  int nd = depth + (is_q ? 1 : 0);
  for (const Regexp* s : re->sub) if (nested_quantifiers(s, nd)) return true;
  return false;
}

static bool can_reach_capture(const Prog& p, int start_idx) {  // analysis.go:408-442
  std::set<int> visited;
  std::deque<int> queue = {start_idx};
  while (!queue.empty()) {
    int curr = queue.front(); queue.pop_front();
    if (curr < 0 || curr >= (int)p.inst.size() || visited.count(curr)) continue;
    visited.insert(curr);
    const Inst& inst = p.inst[curr];
    if (inst.op == InstCapture) return true;
    if (inst.op == InstMatch || inst.op == InstFail) continue;
    if (inst.op == InstAlt) { queue.push_back((int)inst.out); queue.push_back((int)inst.arg); }
    else queue.push_back((int)inst.out);
  }
  return false;
}

static uint64_t epsilon_closure(const Prog& p, int start) {  // analysis.go:462-501
  if (start >= 64) return 0;
  uint64_t result = 0;
  std::set<int> visited;
  std::deque<int> queue = {start};
  while (!queue.empty()) {
    int state = queue.front(); queue.pop_front();
    if (visited.count(state)) continue;
    visited.insert(state);
    if (state < 64) result |= (1ull << state);
    if (state >= (int)p.inst.size()) continue;
    const Inst& inst = p.inst[state];
    switch (inst.op) {
      case InstNop: case InstCapture: queue.push_back((int)inst.out); break;
      case InstAlt: queue.push_back((int)inst.out); queue.push_back((int)inst.arg); break;
    }
  }
  return result;
}

// ---------------------------------------------------------------------------------------
// analysis_match_len.go
static int rune_len(int32_t r) {  // utf8.RuneLen
  if (r < 0) return -1;
  if (r <= 0x7F) return 1;
  if (r <= 0x7FF) return 2;
  if (r >= 0xD800 && r <= 0xDFFF) return -1;
  if (r <= 0xFFFF) return 3;
  if (r <= 0x10FFFF) return 4;
  return -1;
}

static int min_match_len(const Regexp* re) {  // analysis_match_len.go:34-138
  if (re == nullptr) return 0;
  switch (re->op) {
    case OpLiteral: { int t = 0; for (int32_t r : re->rune) t += rune_len(r); return t; }
    case OpCharClass: {
      if (re->rune.empty()) return 0;
      int mn = 4;
      for (size_t i = 0; i + 1 < re->rune.size(); i += 2) mn = std::min(mn, rune_len(re->rune[i]));
      return mn;
    }
    case OpAnyCharNotNL: case OpAnyChar: return 1;
    case OpCapture: return re->sub.empty() ? 0 : min_match_len(re->sub[0]);
    case OpPlus: return re->sub.empty() ? 0 : min_match_len(re->sub[0]);
    case OpRepeat: return re->sub.empty() ? 0 : re->min * min_match_len(re->sub[0]);
    case OpConcat: { int t = 0; for (const Regexp* s : re->sub) t += min_match_len(s); return t; }
    case OpAlternate: {
      if (re->sub.empty()) return 0;
      int mn = min_match_len(re->sub[0]);
      for (size_t i = 1; i < re->sub.size(); i++) mn = std::min(mn, min_match_len(re->sub[i]));
      return mn;
    }
    default: return 0;
  }
}

static int max_match_len(const Regexp* re) {  // analysis_match_len.go:142-251
  if (re == nullptr) return 0;
  switch (re->op) {
    case OpLiteral: { int t = 0; for (int32_t r : re->rune) t += rune_len(r); return t; }
    case OpCharClass: {
      if (re->rune.empty()) return 0;
      int mx = 1;
      for (size_t i = 0; i + 1 < re->rune.size(); i += 2) mx = std::max(mx, rune_len(re->rune[i + 1]));
      return mx;
    }
    case OpAnyCharNotNL: case OpAnyChar: return 4;
    case OpCapture: return re->sub.empty() ? 0 : max_match_len(re->sub[0]);
    case OpStar: case OpPlus: return -1;
    case OpQuest: return re->sub.empty() ? 0 : max_match_len(re->sub[0]);
    case OpRepeat: {
      if (re->max == -1) return -1;
      if (re->sub.empty()) return 0;
      int s = max_match_len(re->sub[0]);
      if (s == -1) return -1;
      return re->max * s;
    }
    case OpConcat: {
      int t = 0;
      for (const Regexp* s : re->sub) { int m = max_match_len(s); if (m == -1) return -1; t += m; }
      return t;
    }
    case OpAlternate: {
      int mx = 0;
      for (const Regexp* s : re->sub) { int m = max_match_len(s); if (m == -1) return -1; mx = std::max(mx, m); }
      return mx;
    }
    default: return 0;
  }
}

// ---------------------------------------------------------------------------------------
// tdfa.go
namespace {

struct NfaState { int id; std::vector<TagAction> actions; };

struct TdfaBuilder {
  const Prog& prog;
  int max_states;
  std::vector<std::vector<NfaState>> states;
  std::map<std::string, int> state_map;
  std::vector<std::map<int, int>> transitions;
  std::vector<std::map<int, std::vector<TagAction>>> tag_actions;
  std::vector<uint8_t> is_accept;
  int start_begin = 0, start_any = 0;
  std::vector<TagAction> init_begin, init_any;

  TdfaBuilder(const Prog& p, int ms) : prog(p), max_states(ms) {}

  static std::vector<TagAction> compact(const std::vector<TagAction>& actions) {  // tdfa.go:426-445
    std::map<int, TagAction> last;
    for (const TagAction& a : actions) last[a.tag] = a;
    std::vector<TagAction> r;
    for (auto& kv : last) r.push_back(kv.second);
    return r;
  }

  std::vector<NfaState> closure(const std::vector<NfaState>& in, bool collect_start, uint32_t match_flags) {
    // tdfa.go:449-510 epsilonClosureWithCaptures
    std::set<int> visited;
    std::vector<NfaState> result;
    std::vector<NfaState> stack(in.rbegin(), in.rend());
    while (!stack.empty()) {
      NfaState st = stack.back(); stack.pop_back();
      st.actions = compact(st.actions);
      if (visited.count(st.id) || st.id >= (int)prog.inst.size()) continue;
      visited.insert(st.id);
      result.push_back(st);
      const Inst& inst = prog.inst[st.id];
      switch (inst.op) {
        case InstNop:
          stack.push_back({(int)inst.out, st.actions});
          break;
        case InstCapture: {
          int tag = (int)inst.arg;
          bool is_start = tag % 2 == 0;
          std::vector<TagAction> na = st.actions;
          if (!is_start || collect_start) na.push_back({tag, 0});
          stack.push_back({(int)inst.out, na});
          break;
        }
        case InstAlt:
          stack.push_back({(int)inst.arg, st.actions});
          stack.push_back({(int)inst.out, st.actions});
          break;
        case InstEmptyWidth:
          if ((inst.arg & match_flags) == inst.arg) stack.push_back({(int)inst.out, st.actions});
          break;
      }
    }
    return result;
  }

  static std::string key(const std::vector<NfaState>& states) {  // tdfa.go:513-539 nfaSetKey
    std::vector<const NfaState*> sorted;
    for (const NfaState& s : states) sorted.push_back(&s);
    std::stable_sort(sorted.begin(), sorted.end(), [](const NfaState* a, const NfaState* b) { return a->id < b->id; });
    std::string k;
    for (size_t i = 0; i < sorted.size(); i++) {
      if (i > 0) k += ",";
      k += std::to_string(sorted[i]->id);
      if (!sorted[i]->actions.empty()) {
        k += "[";
        for (size_t j = 0; j < sorted[i]->actions.size(); j++) {
          if (j > 0) k += ";";
          k += std::to_string(sorted[i]->actions[j].tag) + ":" + std::to_string(sorted[i]->actions[j].offset);
        }
        k += "]";
      }
    }
    return k;
  }

  std::vector<int> possible_chars(const std::vector<NfaState>& set) {  // tdfa.go:293-337
    bool cs[128] = {false};
    for (const NfaState& st : set) {
      const Inst& inst = prog.inst[st.id];
      switch (inst.op) {
        case InstRune1:
          if (!inst.rune.empty() && inst.rune[0] < 128) cs[inst.rune[0]] = true;
          break;
        case InstRune:
          for (size_t i = 0; i + 1 < inst.rune.size(); i += 2) {
            int32_t lo = inst.rune[i], hi = inst.rune[i + 1];
            if (lo < 128) {
              int32_t end = hi >= 128 ? 127 : hi;
              for (int32_t c = lo; c <= end; c++) cs[c] = true;
            }
          }
          break;
        case InstRuneAny:
          for (int c = 0; c < 128; c++) cs[c] = true;
          break;
        case InstRuneAnyNotNL:
          for (int c = 0; c < 128; c++) if (c != '\n') cs[c] = true;
          break;
      }
    }
    std::vector<int> r;
    for (int c = 0; c < 128; c++) if (cs[c]) r.push_back(c);
    return r;
  }

  void transition(const std::vector<NfaState>& set, int c, std::vector<NfaState>& next, std::vector<TagAction>& common) {
    // tdfa.go:340-410 computeTransition
    std::vector<NfaState> ns;
    for (const NfaState& st : set) {
      const Inst& inst = prog.inst[st.id];
      bool m = false;
      switch (inst.op) {
        case InstRune1:
          if (!inst.rune.empty()) { int32_t r = inst.rune[0]; if (r < 128 && r == c) m = true; }
          break;
        case InstRune:
          for (size_t i = 0; i + 1 < inst.rune.size(); i += 2)
            if (c >= inst.rune[i] && c <= inst.rune[i + 1]) { m = true; break; }
          break;
        case InstRuneAny: m = true; break;
        case InstRuneAnyNotNL: m = c != '\n'; break;
      }
      if (m) {
        std::vector<TagAction> pa(st.actions.size());
        for (size_t k = 0; k < st.actions.size(); k++) pa[k] = {st.actions[k].tag, st.actions[k].offset + 1};
        ns.push_back({(int)inst.out, pa});
      }
    }
    next.clear(); common.clear();
    if (ns.empty()) return;
    next = closure(ns, true, 0);
    if (next.empty()) return;
    common = next[0].actions;
    for (size_t i = 1; i < next.size(); i++) {
      size_t ml = std::min(common.size(), next[i].actions.size()), k = 0;
      while (k < ml && common[k] == next[i].actions[k]) k++;
      common.resize(k);
      if (common.empty()) break;
    }
    if (!common.empty())
      for (NfaState& s : next) s.actions.erase(s.actions.begin(), s.actions.begin() + common.size());
  }

  bool set_has_match(const std::vector<NfaState>& set) const {
    for (const NfaState& s : set) if (prog.inst[s.id].op == InstMatch) return true;
    return false;
  }

  int add_state(const std::vector<NfaState>& set, const std::string& k) {
    int idx = (int)states.size();
    states.push_back(set);
    state_map[k] = idx;
    transitions.emplace_back();
    tag_actions.emplace_back();
    is_accept.push_back(set_has_match(set) ? 1 : 0);
    return idx;
  }

  bool build(std::string& err) {  // tdfa.go:111-249
    std::vector<NfaState> start_nfa = {{prog.start, {}}};
    std::vector<NfaState> sb = closure(start_nfa, true, EmptyBeginText);
    if (!sb.empty()) init_begin = sb[0].actions;
    add_state(sb, key(sb));
    start_begin = 0;
    std::vector<NfaState> sa = closure(start_nfa, true, 0);
    if (!sa.empty()) init_any = sa[0].actions;
    std::string ka = key(sa);
    auto it = state_map.find(ka);
    if (it != state_map.end()) start_any = it->second;
    else start_any = add_state(sa, ka);

    std::deque<int> worklist = {0};
    if (start_any != 0) worklist.push_back(start_any);
    std::set<int> processed;
    while (!worklist.empty()) {
      int si = worklist.front(); worklist.pop_front();
      if (processed.count(si)) continue;
      processed.insert(si);
      std::vector<NfaState> cur = states[si];  // copy: states may reallocate
      for (int c : possible_chars(cur)) {
        std::vector<NfaState> next; std::vector<TagAction> actions;
        transition(cur, c, next, actions);
        if (next.empty()) continue;
        std::string k = key(next);
        int ni;
        auto f = state_map.find(k);
        if (f == state_map.end()) {
          ni = (int)states.size();
          if (ni >= max_states) { err = "TDFA state explosion"; return false; }
          add_state(next, k);
          worklist.push_back(ni);
        } else {
          ni = f->second;
        }
        transitions[si][c] = ni;
        if (!actions.empty()) tag_actions[si][c] = actions;
      }
    }
    return true;
  }
};

}  // namespace

static bool build_tdfa(const Program& P, Tdfa& t, int max_states) {
  // tdfa.go:83-108 CanUseTDFA
  for (const Inst& i : P.prog.inst)
    if (i.op == InstEmptyWidth && i.arg != EmptyBeginText && i.arg != EmptyEndText) return false;
  TdfaBuilder b(P.prog, max_states);
  std::string err;
  if (!b.build(err)) return false;
  if ((int)b.states.size() > max_states) return false;
  int ns = (int)b.states.size();
  t.built = true;
  t.num_states = ns;
  int groups = (int)P.capture_names.size();
  if (groups == 0) groups = 1;
  t.num_tags = groups * 2;  // tdfa.go:797-803
  t.start_begin = b.start_begin; t.start_any = b.start_any;
  for (const TagAction& a : b.init_begin) t.init_tags_begin.push_back(a.tag);
  for (const TagAction& a : b.init_any) t.init_tags_any.push_back(a.tag);
  t.trans.assign((size_t)ns * 128, -1);
  t.actions.assign((size_t)ns * 128, {});
  for (int s = 0; s < ns; s++) {
    for (auto& kv : b.transitions[s]) t.trans[(size_t)s * 128 + kv.first] = kv.second;
    for (auto& kv : b.tag_actions[s]) {
      t.actions[(size_t)s * 128 + kv.first] = kv.second;
      t.max_actions = std::max(t.max_actions, (int)kv.second.size());
    }
  }
  t.accept.assign(ns, 0); t.accept_eot.assign(ns, 0); t.accept_actions.assign(ns, {});
  std::vector<uint8_t> has_aa(ns, 0);
  for (int s = 0; s < ns; s++) t.accept[s] = b.is_accept[s];
  // tdfa.go:251-267
  for (int s = 0; s < ns; s++) {
    std::vector<NfaState> cl = b.closure(b.states[s], true, EmptyEndText);
    for (const NfaState& st : cl) {
      if (P.prog.inst[st.id].op == InstMatch) {
        t.accept_eot[s] = 1;
        if (!st.actions.empty()) { t.accept_actions[s] = TdfaBuilder::compact(st.actions); has_aa[s] = 1; }
        break;
      }
    }
  }
  // tdfa.go:269-287
  for (int s = 0; s < ns; s++) {
    if (!t.accept[s]) continue;
    if (has_aa[s]) continue;
    for (const NfaState& st : b.states[s]) {
      if (P.prog.inst[st.id].op == InstMatch) {
        if (!st.actions.empty()) t.accept_actions[s] = TdfaBuilder::compact(st.actions);
        break;
      }
    }
  }
  for (int s = 0; s < ns; s++) t.max_accept_actions = std::max(t.max_accept_actions, (int)t.accept_actions[s].size());
  return true;
}

// ---------------------------------------------------------------------------------------
static bool ast_has(const Regexp* re, bool (*pred)(const Regexp*)) {
  if (re == nullptr) return false;
  if (pred(re)) return true;
  for (const Regexp* s : re->sub) if (ast_has(s, pred)) return true;
  return false;
}

static void derive_labels(Program& P, const Regexp* ast) {
  // analyze_api.go:76-134 deriveFeatureLabels
  std::vector<std::string>& fl = P.feature_labels;
  const std::string& pat = P.pattern;
  if (P.anchored || P.end_anchor) fl.push_back("Anchored");
  if (ast_has(ast, [](const Regexp* r) { return r->op == OpAlternate; })) fl.push_back("Alternation");
  if (P.prog.num_cap > 2) fl.push_back("Captures");
  bool cc = pat.find_first_of("[]") != std::string::npos;
  for (const char* e : {"\\d", "\\D", "\\w", "\\W", "\\s", "\\S"}) if (pat.find(e) != std::string::npos) cc = true;
  if (!cc) cc = ast_has(ast, [](const Regexp* r) { return r->op == OpCharClass || r->op == OpAnyCharNotNL || r->op == OpAnyChar; });
  if (cc) fl.push_back("CharClass");
  bool mb = false;
  for (unsigned char c : pat) if (c > 127) mb = true;
  if (mb) fl.push_back("Multibyte");
  if (pat.find("(?:") != std::string::npos) fl.push_back("NonCapturing");
  if (ast_has(ast, [](const Regexp* r) { return r->op == OpStar || r->op == OpPlus || r->op == OpQuest || r->op == OpRepeat; }))
    fl.push_back("Quantifiers");
  if (P.has_word_boundary) fl.push_back("WordBoundary");
  if (has_unicode_char_class(P.prog)) fl.push_back("UnicodeCharClass");
  if (fl.empty()) fl.push_back("Simple");
  std::sort(fl.begin(), fl.end());

  // analyze_api.go:137-186 deriveEngineLabels (+ :190-218 canUseTDFAStandalone heuristic)
  std::vector<std::string>& el = P.engine_labels;
  bool thompson = P.use_thompson_nfa;
  bool memo = P.catastrophic_risk && !thompson;
  bool tdfa = false, tnfa = false;
  if (P.has_captures && P.catastrophic_risk) {
    int thr = P.opts.tdfa_threshold > 0 ? P.opts.tdfa_threshold : 500;
    bool ok = true;
    for (const Inst& i : P.prog.inst)
      if (i.op == InstEmptyWidth && (i.arg & (EmptyWordBoundary | EmptyNoWordBoundary))) ok = false;
    if ((int)P.prog.inst.size() > thr / 10) ok = false;
    if (ok) tdfa = true; else tnfa = true;
  }
  if (thompson) el.push_back("Thompson");
  if (P.has_captures) { if (tdfa) el.push_back("TDFA"); else if (tnfa) el.push_back("TNFA"); }
  if (memo && !thompson) el.push_back("Memoization");
  if (!thompson && !tdfa && !tnfa && !memo) el.push_back("Backtracking");
  std::sort(el.begin(), el.end());
}

bool build_program(const std::string& pattern, const Options& opts, Program& P, std::string& err) {
  P = Program();
  P.pattern = pattern;
  P.opts = opts;
  Arena arena;
  Regexp* ast = nullptr;
  if (!parse(pattern, Perl, arena, &ast, err)) return false;   // regengo.go:92
  ast = simplify(ast, arena);                                   // regengo.go:98 (:101 is a no-op)
  if (!compile(ast, P.prog, err)) return false;                 // regengo.go:104
  const Prog& prog = P.prog;
  size_t n = prog.inst.size();

  P.has_captures = prog.num_cap > 2;                            // regengo.go:110
  {
    std::map<int, std::string> cap_map; int max_cap = 0;
    walk_capture_names(ast, cap_map, max_cap);
    P.capture_names.assign(max_cap + 1, "");
    for (auto& kv : cap_map) P.capture_names[kv.first] = kv.second;
    if (!P.has_captures) P.capture_names.clear();               // compiler.go:67-70 (only WithCaptures)
  }
  // compiler.go:73-84
  P.needs_backtracking = needs_backtracking(prog);
  P.anchored = is_anchored(prog);
  bool memo0 = detect_complexity(prog);
  P.alt_ckpt.assign(n, 0);
  int n_ckpt = 0;
  for (size_t i = 0; i < n; i++)
    if (prog.inst[i].op == InstAlt && can_reach_capture(prog, (int)prog.inst[i].out)) { P.alt_ckpt[i] = 1; n_ckpt++; }
  P.per_capture_ckpt = n_ckpt > 3;                              // analysis.go:396-404
  P.has_word_boundary = has_word_boundary(prog);
  // analysis.go:282-314
  P.end_anchor = has_end_anchor(prog);
  P.nested_loops = memo0;
  P.catastrophic_risk = nested_quantifiers(ast, 0);
  P.use_thompson_nfa = (P.catastrophic_risk || P.nested_loops) && !P.end_anchor;
  // match length (compiler.go:105-113), stream defaults (streaming.go:25-62, 87-96)
  P.min_match_len = min_match_len(ast);
  P.max_match_len = max_match_len(ast);
  P.default_max_leftover = 1 << 20;
  if (P.max_match_len != -1) {
    P.default_max_leftover = P.max_match_len * 10;
    if (P.default_max_leftover < 1024) P.default_max_leftover = 1024;
    if (P.default_max_leftover > (1 << 20)) P.default_max_leftover = 1 << 20;
  }
  P.min_buffer = 64 * 1024;
  if (P.max_match_len > 0) P.min_buffer = std::max(P.max_match_len * 2, 64 * 1024);

  // compiler.go:127-153 engine selection
  bool use_thompson_for_match = opts.force_thompson || P.use_thompson_nfa;
  bool memo = memo0;
  if (P.catastrophic_risk && !use_thompson_for_match) memo = true;
  bool use_tdfa = false, use_tnfa = false;
  if (P.has_captures && (P.catastrophic_risk || opts.force_tdfa)) {
    int thr = opts.tdfa_threshold > 0 ? opts.tdfa_threshold : 500;
    Tdfa t;
    if (build_tdfa(P, t, thr)) { use_tdfa = true; P.tdfa = t; }
    else use_tnfa = true;
  } else if (opts.force_tnfa) {
    use_tnfa = true;
  }
  // compiler.go:262-295: Thompson only when <= 64 instructions, else backtracking (Q11)
  P.match_engine = (use_thompson_for_match && n <= 64) ? MATCH_THOMPSON : MATCH_BT;
  P.match_memo = memo;
  if (P.has_captures) {
    if (use_tdfa) { P.find_engine = FIND_TDFA; P.find_memo = false; }
    else if (use_tnfa) { P.find_engine = FIND_BT; P.find_memo = true; P.tnfa = true; }  // compiler.go:423-426
    else { P.find_engine = FIND_BT; P.find_memo = memo; }
  }

  // findRequiredPrefix (compiler.go:719-737)
  {
    int pc = prog.start;
    size_t guard = 0;
    while (guard++ <= n) {
      const Inst& in = prog.inst[pc];
      if (in.op == InstNop || in.op == InstCapture) { pc = (int)in.out; continue; }
      if (in.op == InstRune1 && in.rune.size() == 1 && in.rune[0] < 128) { P.has_prefix = true; P.prefix = (uint8_t)in.rune[0]; }
      break;
    }
  }
  // Match-mode simple greedy loops (instructions.go:414, 462-479)
  P.greedy_loop.assign(n, 0);
  for (size_t i = 0; i < n; i++) {
    const Inst& in = prog.inst[i];
    if (in.op != InstAlt || !(in.out < i)) continue;
    uint8_t t = prog.inst[in.out].op;
    if (t == InstRune || t == InstRune1 || t == InstRuneAny || t == InstRuneAnyNotNL) P.greedy_loop[i] = 1;
  }
  // class bitmaps (charclass.go:10-54, instructions.go:205-249)
  P.unicode_class.assign(n, 0);
  P.class_bits.assign(n * 8, 0);
  for (size_t i = 0; i < n; i++) {
    const Inst& in = prog.inst[i];
    if (in.op != InstRune) continue;
    bool ascii_only = true;
    for (size_t k = 0; k + 1 < in.rune.size(); k += 2) if (in.rune[k + 1] >= 128) ascii_only = false;
    P.unicode_class[i] = ascii_only ? 0 : 1;
    for (size_t k = 0; k + 1 < in.rune.size(); k += 2) {
      int32_t lo = in.rune[k], hi = in.rune[k + 1];
      if (lo >= 128) continue;
      if (hi >= 128) hi = 127;
      for (int32_t c = lo; c <= hi; c++) P.class_bits[i * 8 + (c >> 5)] |= (1u << (c & 31));
    }
  }
  // Thompson constants (thompson.go:25-60)
  P.closures.assign(n, 0); P.eps_after.assign(n, 0); P.char_state.assign(n, 0);
  for (size_t i = 0; i < n; i++) P.closures[i] = epsilon_closure(prog, (int)i);
  if (prog.start < (int)n) P.start_closure = P.closures[prog.start];
  for (size_t i = 0; i < n; i++) {
    const Inst& in = prog.inst[i];
    if (in.op == InstMatch && i < 64) P.accept_mask |= (1ull << i);
    if (in.op == InstRune || in.op == InstRune1 || in.op == InstRuneAny || in.op == InstRuneAnyNotNL) {
      P.char_state[i] = 1;
      if (in.out < n) P.eps_after[i] = P.closures[in.out];
    }
  }
  // Thompson per-state byte conditions (thompson.go:199-303): the condition is on the BYTE c.
  P.thompson_cond.assign(n * 8, 0);
  for (size_t i = 0; i < n; i++) {
    const Inst& in = prog.inst[i];
    auto set = [&](uint32_t c) { P.thompson_cond[i * 8 + ((c & 255) >> 5)] |= 1u << (c & 31); };
    if (in.op == InstRune1) {
      if (!in.rune.empty()) set((uint32_t)in.rune[0] & 255u);            // byte(r): truncated (Q9)
    } else if (in.op == InstRuneAny) {
      for (uint32_t c = 0; c < 256; c++) set(c);
    } else if (in.op == InstRuneAnyNotNL) {
      for (uint32_t c = 0; c < 256; c++) if (c != '\n') set(c);
    } else if (in.op == InstRune) {
      const auto& r = in.rune;
      bool fold = (in.arg & FoldCase) != 0;
      if (r.size() == 2 && r[0] == r[1]) {
        if (fold && r[0] < 128) { for (uint32_t c = 0; c < 256; c++) if ((c | 0x20u) == (uint32_t)(r[0] | 0x20)) set(c); }
        else set((uint32_t)r[0] & 255u);
      } else {
        for (size_t k = 0; k + 1 < r.size(); k += 2) {
          int32_t lo = r[k], hi = r[k + 1];
          if (lo >= 128) continue;
          if (lo == hi) { set((uint32_t)lo); continue; }
          if (hi > 127) hi = 127;
          for (int32_t c = lo; c <= hi; c++) set((uint32_t)c);
        }
      }
    }
  }
  derive_labels(P, ast);
  return true;
}

// ---------------------------------------------------------------------------------------
static void jstr(std::ostringstream& o, const std::string& s) {
  o << '"';
  for (unsigned char c : s) {
    if (c == '"' || c == '\\') o << '\\' << c;
    else if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o << b; }
    else o << c;
  }
  o << '"';
}
template <class T> static void jarr(std::ostringstream& o, const std::vector<T>& v) {
  o << '[';
  for (size_t i = 0; i < v.size(); i++) { if (i) o << ','; o << (long long)v[i]; }
  o << ']';
}
static void jacts(std::ostringstream& o, const std::vector<TagAction>& v) {
  o << '[';
  for (size_t i = 0; i < v.size(); i++) { if (i) o << ','; o << '[' << v[i].tag << ',' << v[i].offset << ']'; }
  o << ']';
}

std::string program_to_json(const Program& P) {
  std::ostringstream o;
  o << "{\"pattern\":"; jstr(o, P.pattern);
  o << ",\"start\":" << P.prog.start << ",\"num_cap\":" << P.prog.num_cap << ",\"inst\":[";
  for (size_t i = 0; i < P.prog.inst.size(); i++) {
    const Inst& in = P.prog.inst[i];
    if (i) o << ',';
    o << "{\"op\":" << (int)in.op << ",\"out\":" << in.out << ",\"arg\":" << in.arg << ",\"rune\":";
    jarr(o, in.rune);
    o << '}';
  }
  o << "],\"capture_names\":[";
  for (size_t i = 0; i < P.capture_names.size(); i++) { if (i) o << ','; jstr(o, P.capture_names[i]); }
  o << "],\"has_captures\":" << P.has_captures << ",\"needs_backtracking\":" << P.needs_backtracking
    << ",\"anchored\":" << P.anchored << ",\"has_word_boundary\":" << P.has_word_boundary
    << ",\"nested_loops\":" << P.nested_loops << ",\"catastrophic_risk\":" << P.catastrophic_risk
    << ",\"end_anchor\":" << P.end_anchor << ",\"use_thompson_nfa\":" << P.use_thompson_nfa
    << ",\"match_engine\":" << P.match_engine << ",\"match_memo\":" << P.match_memo
    << ",\"find_engine\":" << P.find_engine << ",\"find_memo\":" << P.find_memo << ",\"tnfa\":" << P.tnfa
    << ",\"per_capture_ckpt\":" << P.per_capture_ckpt << ",\"has_prefix\":" << P.has_prefix
    << ",\"prefix\":" << (int)P.prefix << ",\"min_match_len\":" << P.min_match_len
    << ",\"max_match_len\":" << P.max_match_len << ",\"default_max_leftover\":" << P.default_max_leftover
    << ",\"min_buffer\":" << P.min_buffer;
  o << ",\"alt_ckpt\":"; jarr(o, P.alt_ckpt);
  o << ",\"greedy_loop\":"; jarr(o, P.greedy_loop);
  o << ",\"unicode_class\":"; jarr(o, P.unicode_class);
  o << ",\"class_bits\":"; jarr(o, P.class_bits);
  // uint64 masks as decimal strings (JSON numbers lose precision above 2^53)
  o << ",\"start_closure\":\"" << P.start_closure << "\",\"accept_mask\":\"" << P.accept_mask << "\"";
  o << ",\"closures\":[";
  for (size_t i = 0; i < P.closures.size(); i++) { if (i) o << ','; o << '"' << P.closures[i] << '"'; }
  o << "],\"eps_after\":[";
  for (size_t i = 0; i < P.eps_after.size(); i++) { if (i) o << ','; o << '"' << P.eps_after[i] << '"'; }
  o << "],\"char_state\":"; jarr(o, P.char_state);
  o << ",\"thompson_cond\":"; jarr(o, P.thompson_cond);
  o << ",\"engine_labels\":[";
  for (size_t i = 0; i < P.engine_labels.size(); i++) { if (i) o << ','; jstr(o, P.engine_labels[i]); }
  o << "],\"feature_labels\":[";
  for (size_t i = 0; i < P.feature_labels.size(); i++) { if (i) o << ','; jstr(o, P.feature_labels[i]); }
  o << "]";
  const Tdfa& t = P.tdfa;
  o << ",\"tdfa\":";
  if (!t.built) {
    o << "null";
  } else {
    o << "{\"num_states\":" << t.num_states << ",\"num_tags\":" << t.num_tags << ",\"start_begin\":" << t.start_begin
      << ",\"start_any\":" << t.start_any << ",\"max_actions\":" << t.max_actions
      << ",\"max_accept_actions\":" << t.max_accept_actions;
    o << ",\"init_tags_begin\":"; jarr(o, t.init_tags_begin);
    o << ",\"init_tags_any\":"; jarr(o, t.init_tags_any);
    o << ",\"trans\":"; jarr(o, t.trans);
    o << ",\"accept\":"; jarr(o, t.accept);
    o << ",\"accept_eot\":"; jarr(o, t.accept_eot);
    o << ",\"actions\":{";
    bool first = true;
    for (size_t i = 0; i < t.actions.size(); i++) {
      if (t.actions[i].empty()) continue;
      if (!first) o << ',';
      first = false;
      o << '"' << i << "\":"; jacts(o, t.actions[i]);
    }
    o << "},\"accept_actions\":[";
    for (size_t i = 0; i < t.accept_actions.size(); i++) { if (i) o << ','; jacts(o, t.accept_actions[i]); }
    o << "]}";
  }
  o << "}";
  return o.str();
}

}  // namespace rgx
