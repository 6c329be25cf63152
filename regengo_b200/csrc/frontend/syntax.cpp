// Restatement of Go regexp/syntax: parse.go (Parse with syntax.Perl), simplify.go, compile.go.
// See syntax.hpp for why this exists.  Call sites in the reference: regengo.go:92 (Parse),
// regengo.go:98 (Simplify), regengo.go:104 (Compile); same three calls in
// internal/compiler/analyze_api.go:32-41.
//
// Deliberately unsupported (load fails loudly, no fallback): \p{..}/\P{..} Unicode groups,
// case folding ((?i), or a class such as [Aa] that Go rewrites into a folded literal -- the
// reference generator itself indexes out of range on that), \C.
#include "syntax.hpp"

#include <algorithm>
#include <cstring>

namespace rgx {

static const int32_t MaxRune = 0x10FFFF;

// ---------------------------------------------------------------------------------------
// UTF-8 decode of the pattern text (Go: nextRune / utf8.DecodeRuneInString).
static bool next_rune(const std::string& s, size_t& pos, int32_t& r, std::string& err) {
  unsigned char c0 = (unsigned char)s[pos];
  if (c0 < 0x80) { r = c0; pos += 1; return true; }
  int need = 0; int32_t v = 0; int32_t minv = 0;
  if ((c0 & 0xE0) == 0xC0) { need = 1; v = c0 & 0x1F; minv = 0x80; }
  else if ((c0 & 0xF0) == 0xE0) { need = 2; v = c0 & 0x0F; minv = 0x800; }
  else if ((c0 & 0xF8) == 0xF0) { need = 3; v = c0 & 0x07; minv = 0x10000; }
  else { err = "invalid UTF-8"; return false; }
  if (pos + need >= s.size()) { err = "invalid UTF-8"; return false; }
  for (int i = 1; i <= need; i++) {
    unsigned char c = (unsigned char)s[pos + i];
    if ((c & 0xC0) != 0x80) { err = "invalid UTF-8"; return false; }
    v = (v << 6) | (c & 0x3F);
  }
  if (v < minv || v > MaxRune || (v >= 0xD800 && v <= 0xDFFF)) { err = "invalid UTF-8"; return false; }
  r = v; pos += need + 1; return true;
}

// unicode.SimpleFold restricted to what the supported subset needs: ASCII letters, with the
// two ASCII orbits that leave ASCII (K -> k -> U+212A, S -> s -> U+017F).  Other runes fold
// to themselves here, which only makes the parser *less* eager to create folded literals.
static int32_t simple_fold(int32_t r) {
  if (r >= 'A' && r <= 'Z') return r + 32;
  if (r == 'k') return 0x212A;
  if (r == 's') return 0x017F;
  if (r >= 'a' && r <= 'z') return r - 32;
  if (r == 0x212A) return 'K';
  if (r == 0x017F) return 'S';
  return r;
}

// ---------------------------------------------------------------------------------------
// Character-class helpers (parse.go: appendRange, cleanClass, negateClass, ...).
static void append_range(std::vector<int32_t>& r, int32_t lo, int32_t hi) {
  size_t n = r.size();
  for (size_t i = 2; i <= 4; i += 2) {
    if (n >= i) {
      int32_t rlo = r[n - i], rhi = r[n - i + 1];
      if (lo <= rhi + 1 && rlo <= hi + 1) {
        if (lo < rlo) r[n - i] = lo;
        if (hi > rhi) r[n - i + 1] = hi;
        return;
      }
    }
  }
  r.push_back(lo); r.push_back(hi);
}

static void append_class(std::vector<int32_t>& r, const std::vector<int32_t>& x) {
  for (size_t i = 0; i + 1 < x.size(); i += 2) append_range(r, x[i], x[i + 1]);
}

static void append_negated_class(std::vector<int32_t>& r, const std::vector<int32_t>& x) {
  int32_t next_lo = 0;
  for (size_t i = 0; i + 1 < x.size(); i += 2) {
    int32_t lo = x[i], hi = x[i + 1];
    if (next_lo <= lo - 1) append_range(r, next_lo, lo - 1);
    next_lo = hi + 1;
  }
  if (next_lo <= MaxRune) append_range(r, next_lo, MaxRune);
}

static void append_literal(std::vector<int32_t>& r, int32_t x, uint32_t /*flags*/) {
  // FoldCase literals are rejected before they can reach here.
  append_range(r, x, x);
}

static void clean_class(std::vector<int32_t>& r) {
  // Sort by lo increasing, hi decreasing to break ties.
  size_t np = r.size() / 2;
  std::vector<std::pair<int32_t, int32_t>> p(np);
  for (size_t i = 0; i < np; i++) p[i] = {r[2 * i], r[2 * i + 1]};
  std::sort(p.begin(), p.end(), [](const std::pair<int32_t, int32_t>& a, const std::pair<int32_t, int32_t>& b) {
    return a.first < b.first || (a.first == b.first && a.second > b.second);
  });
  for (size_t i = 0; i < np; i++) { r[2 * i] = p[i].first; r[2 * i + 1] = p[i].second; }
  if (r.size() < 2) return;
  size_t w = 2;
  for (size_t i = 2; i < r.size(); i += 2) {
    int32_t lo = r[i], hi = r[i + 1];
    if (lo <= r[w - 1] + 1) {
      if (hi > r[w - 1]) r[w - 1] = hi;
      continue;
    }
    r[w] = lo; r[w + 1] = hi; w += 2;
  }
  r.resize(w);
}

static void negate_class(std::vector<int32_t>& r) {
  int32_t next_lo = 0;
  std::vector<int32_t> out;
  for (size_t i = 0; i < r.size(); i += 2) {
    int32_t lo = r[i], hi = r[i + 1];
    if (next_lo <= lo - 1) { out.push_back(next_lo); out.push_back(lo - 1); }
    next_lo = hi + 1;
  }
  if (next_lo <= MaxRune) { out.push_back(next_lo); out.push_back(MaxRune); }
  r.swap(out);
}

#include "unicode_tables.inc"

// parse.go: unicodeTable.  "Any", then unicode.Categories[name], then unicode.Scripts[name] (exact, case-sensitive names).
static bool unicode_table(const std::string& name, std::vector<int32_t>& pairs) {
  pairs.clear();
  if (name == "Any") { pairs = {0, MaxRune}; return true; }
  for (const UniTable& t : kUniTables)
    if (name == t.name) { pairs.assign(t.pairs, t.pairs + 2 * (size_t)t.n_pairs); return true; }
  return false;
}

// unicode.IsLetter / unicode.IsDigit for the replace-template parser (replace_template.cpp)
static bool rune_in_table(const char* name, int32_t r) {
  for (const UniTable& t : kUniTables)
    if (std::string(name) == t.name) {
      int lo = 0, hi = t.n_pairs - 1;
      while (lo <= hi) {
        const int mid = (lo + hi) / 2;
        if (r < t.pairs[2 * mid]) hi = mid - 1;
        else if (r > t.pairs[2 * mid + 1]) lo = mid + 1;
        else return true;
      }
      return false;
    }
  return false;
}
bool rune_is_letter(int32_t r) { return rune_in_table("L", r); }
bool rune_is_digit(int32_t r) { return rune_in_table("Nd", r); }

static const std::vector<int32_t> kPerlD = {'0', '9'};
static const std::vector<int32_t> kPerlS = {'\t', '\n', '\f', '\r', ' ', ' '};
static const std::vector<int32_t> kPerlW = {'0', '9', 'A', 'Z', '_', '_', 'a', 'z'};

struct PosixGroup { const char* name; int sign; std::vector<int32_t> cls; };
static const std::vector<PosixGroup>& posix_groups() {
  static const std::vector<PosixGroup> g = {
      {"[:alnum:]", +1, {'0', '9', 'A', 'Z', 'a', 'z'}},
      {"[:alpha:]", +1, {'A', 'Z', 'a', 'z'}},
      {"[:ascii:]", +1, {0x0, 0x7F}},
      {"[:blank:]", +1, {'\t', '\t', ' ', ' '}},
      {"[:cntrl:]", +1, {0x0, 0x1F, 0x7F, 0x7F}},
      {"[:digit:]", +1, {'0', '9'}},
      {"[:graph:]", +1, {'!', '~'}},
      {"[:lower:]", +1, {'a', 'z'}},
      {"[:print:]", +1, {' ', '~'}},
      {"[:punct:]", +1, {'!', '/', ':', '@', '[', '`', '{', '~'}},
      {"[:space:]", +1, {'\t', '\r', ' ', ' '}},
      {"[:upper:]", +1, {'A', 'Z'}},
      {"[:word:]", +1, {'0', '9', 'A', 'Z', '_', '_', 'a', 'z'}},
      {"[:xdigit:]", +1, {'0', '9', 'A', 'F', 'a', 'f'}},
  };
  return g;
}

static bool is_char_class(const Regexp* re) {
  return (re->op == OpLiteral && re->rune.size() == 1) || re->op == OpCharClass ||
         re->op == OpAnyCharNotNL || re->op == OpAnyChar;
}

static bool match_rune(const Regexp* re, int32_t r) {
  switch (re->op) {
    case OpLiteral: return re->rune.size() == 1 && re->rune[0] == r;
    case OpCharClass:
      for (size_t i = 0; i + 1 < re->rune.size(); i += 2)
        if (re->rune[i] <= r && r <= re->rune[i + 1]) return true;
      return false;
    case OpAnyCharNotNL: return r != '\n';
    case OpAnyChar: return true;
  }
  return false;
}

static void merge_char_class(Regexp* dst, Regexp* src) {
  switch (dst->op) {
    case OpAnyChar: break;
    case OpAnyCharNotNL:
      if (match_rune(src, '\n')) dst->op = OpAnyChar;
      break;
    case OpCharClass:
      if (src->op == OpLiteral) append_literal(dst->rune, src->rune[0], src->flags);
      else append_class(dst->rune, src->rune);
      break;
    case OpLiteral: {
      if (src->rune[0] == dst->rune[0] && src->flags == dst->flags) break;
      dst->op = OpCharClass;
      int32_t d0 = dst->rune[0];
      dst->rune.clear();
      append_literal(dst->rune, d0, dst->flags);
      append_literal(dst->rune, src->rune[0], src->flags);
      break;
    }
  }
}

static void clean_alt(Regexp* re) {
  if (re->op == OpCharClass) {
    clean_class(re->rune);
    if (re->rune.size() == 2 && re->rune[0] == 0 && re->rune[1] == MaxRune) {
      re->rune.clear(); re->op = OpAnyChar; return;
    }
    if (re->rune.size() == 4 && re->rune[0] == 0 && re->rune[1] == '\n' - 1 &&
        re->rune[2] == '\n' + 1 && re->rune[3] == MaxRune) {
      re->rune.clear(); re->op = OpAnyCharNotNL; return;
    }
  }
}

static bool regexp_equal(const Regexp* x, const Regexp* y) {
  if (x == nullptr || y == nullptr) return x == y;
  if (x->op != y->op) return false;
  switch (x->op) {
    case OpEndText:
      if ((x->flags & WasDollar) != (y->flags & WasDollar)) return false;
      break;
    case OpLiteral:
    case OpCharClass:
      if ((x->flags & FoldCase) != (y->flags & FoldCase)) return false;
      return x->rune == y->rune;
    case OpAlternate:
    case OpConcat:
      if (x->sub.size() != y->sub.size()) return false;
      for (size_t i = 0; i < x->sub.size(); i++)
        if (!regexp_equal(x->sub[i], y->sub[i])) return false;
      break;
    case OpStar:
    case OpPlus:
    case OpQuest:
      if ((x->flags & NonGreedy) != (y->flags & NonGreedy) || !regexp_equal(x->sub[0], y->sub[0])) return false;
      break;
    case OpRepeat:
      if ((x->flags & NonGreedy) != (y->flags & NonGreedy) || x->min != y->min || x->max != y->max ||
          !regexp_equal(x->sub[0], y->sub[0]))
        return false;
      break;
    case OpCapture:
      if (x->cap != y->cap || x->name != y->name || !regexp_equal(x->sub[0], y->sub[0])) return false;
      break;
  }
  return true;
}

// ---------------------------------------------------------------------------------------
struct Parser {
  Arena& arena;
  uint32_t flags;
  std::vector<Regexp*> stack;
  int num_cap = 0;
  std::string err;

  explicit Parser(Arena& a, uint32_t f) : arena(a), flags(f) {}

  Regexp* new_regexp(int op) { return arena.make(op); }

  // parse.go: (*parser).maybeConcat
  bool maybe_concat(int32_t r, uint32_t fl) {
    size_t n = stack.size();
    if (n < 2) return false;
    Regexp* re1 = stack[n - 1];
    Regexp* re2 = stack[n - 2];
    if (re1->op != OpLiteral || re2->op != OpLiteral || (re1->flags & FoldCase) != (re2->flags & FoldCase))
      return false;
    re2->rune.insert(re2->rune.end(), re1->rune.begin(), re1->rune.end());
    if (r >= 0) {
      re1->rune.assign(1, r);
      re1->flags = fl;
      return true;
    }
    stack.pop_back();
    return false;
  }

  // parse.go: (*parser).push
  Regexp* push(Regexp* re) {
    if (re->op == OpCharClass && re->rune.size() == 2 && re->rune[0] == re->rune[1]) {
      if (maybe_concat(re->rune[0], flags & ~FoldCase)) return nullptr;
      re->op = OpLiteral;
      re->rune.resize(1);
      re->flags = flags & ~FoldCase;
    } else if ((re->op == OpCharClass && re->rune.size() == 4 && re->rune[0] == re->rune[1] &&
                re->rune[2] == re->rune[3] && simple_fold(re->rune[0]) == re->rune[2] &&
                simple_fold(re->rune[2]) == re->rune[0]) ||
               (re->op == OpCharClass && re->rune.size() == 2 && re->rune[0] + 1 == re->rune[1] &&
                simple_fold(re->rune[0]) == re->rune[1] && simple_fold(re->rune[1]) == re->rune[0])) {
      // Go rewrites [Aa] into a case-folded literal; the reference generator cannot emit it.
      if (err.empty()) err = "unsupported: case-folded literal (class like [Aa])";
      re->op = OpLiteral;
      re->rune.resize(1);
      re->flags = flags | FoldCase;
    } else {
      maybe_concat(-1, 0);
    }
    stack.push_back(re);
    return re;
  }

  void literal(int32_t r) {
    Regexp* re = new_regexp(OpLiteral);
    re->flags = flags;
    re->rune.assign(1, r);
    push(re);
  }

  Regexp* op(int o) {
    Regexp* re = new_regexp(o);
    re->flags = flags;
    return push(re);
  }

  // parse.go: (*parser).repeat
  bool repeat(int o, int mn, int mx, const std::string& s, size_t& pos, bool last_repeat) {
    uint32_t fl = flags;
    if (flags & PerlX) {
      if (pos < s.size() && s[pos] == '?') { pos++; fl ^= NonGreedy; }
      if (last_repeat) { err = "invalid nested repetition operator"; return false; }
    }
    size_t n = stack.size();
    if (n == 0) { err = "missing argument to repetition operator"; return false; }
    Regexp* sub = stack[n - 1];
    if (sub->op >= opPseudo) { err = "missing argument to repetition operator"; return false; }
    Regexp* re = new_regexp(o);
    re->min = mn; re->max = mx; re->flags = fl;
    re->sub.assign(1, sub);
    stack[n - 1] = re;
    if (o == OpRepeat && (mn >= 2 || mx >= 2) && !repeat_is_valid(re, 1000)) {
      err = "invalid repeat count"; return false;
    }
    return true;
  }

  static bool repeat_is_valid(Regexp* re, int n) {
    if (re->op == OpRepeat) {
      int m = re->max;
      if (m == 0) return true;
      if (m < 0) m = re->min;
      if (m > n) return false;
      if (m > 0) n /= m;
    }
    for (Regexp* sub : re->sub)
      if (!repeat_is_valid(sub, n)) return false;
    return true;
  }

  // parse.go: (*parser).concat / alternate / collapse / factor
  Regexp* concat() {
    maybe_concat(-1, 0);
    size_t i = stack.size();
    while (i > 0 && stack[i - 1]->op < opPseudo) i--;
    std::vector<Regexp*> subs(stack.begin() + i, stack.end());
    stack.resize(i);
    if (subs.empty()) return push(new_regexp(OpEmptyMatch));
    return push(collapse(subs, OpConcat));
  }

  Regexp* alternate() {
    size_t i = stack.size();
    while (i > 0 && stack[i - 1]->op < opPseudo) i--;
    std::vector<Regexp*> subs(stack.begin() + i, stack.end());
    stack.resize(i);
    if (!subs.empty()) clean_alt(subs.back());
    if (subs.empty()) return push(new_regexp(OpNoMatch));
    return push(collapse(subs, OpAlternate));
  }

  Regexp* collapse(std::vector<Regexp*>& subs, int o) {
    if (subs.size() == 1) return subs[0];
    Regexp* re = new_regexp(o);
    for (Regexp* sub : subs) {
      if (sub->op == o) re->sub.insert(re->sub.end(), sub->sub.begin(), sub->sub.end());
      else re->sub.push_back(sub);
    }
    if (o == OpAlternate) {
      re->sub = factor(re->sub);
      if (re->sub.size() == 1) re = re->sub[0];
    }
    return re;
  }

  static void leading_string(Regexp* re, const std::vector<int32_t>** str, uint32_t* fl) {
    static const std::vector<int32_t> empty;
    if (re->op == OpConcat && !re->sub.empty()) re = re->sub[0];
    if (re->op != OpLiteral) { *str = &empty; *fl = 0; return; }
    *str = &re->rune; *fl = re->flags & FoldCase;
  }

  Regexp* remove_leading_string(Regexp* re, size_t n) {
    if (re->op == OpConcat && !re->sub.empty()) {
      Regexp* sub = remove_leading_string(re->sub[0], n);
      re->sub[0] = sub;
      if (sub->op == OpEmptyMatch) {
        switch (re->sub.size()) {
          case 0:
          case 1:
            re->op = OpEmptyMatch; re->sub.clear(); break;
          case 2:
            re = re->sub[1]; break;
          default:
            re->sub.erase(re->sub.begin()); break;
        }
      }
      return re;
    }
    if (re->op == OpLiteral) {
      re->rune.erase(re->rune.begin(), re->rune.begin() + std::min(n, re->rune.size()));
      if (re->rune.empty()) re->op = OpEmptyMatch;
    }
    return re;
  }

  static Regexp* leading_regexp(Regexp* re) {
    if (re->op == OpEmptyMatch) return nullptr;
    if (re->op == OpConcat && !re->sub.empty()) {
      Regexp* sub = re->sub[0];
      if (sub->op == OpEmptyMatch) return nullptr;
      return sub;
    }
    return re;
  }

  Regexp* remove_leading_regexp(Regexp* re) {
    if (re->op == OpConcat && !re->sub.empty()) {
      re->sub.erase(re->sub.begin());
      switch (re->sub.size()) {
        case 0: re->op = OpEmptyMatch; re->sub.clear(); break;
        case 1: re = re->sub[0]; break;
      }
      return re;
    }
    return new_regexp(OpEmptyMatch);
  }

  std::vector<Regexp*> factor(std::vector<Regexp*> sub) {
    if (sub.size() < 2) return sub;

    // Round 1: Factor out common literal prefixes.
    {
      std::vector<int32_t> str;
      uint32_t strflags = 0;
      size_t start = 0;
      std::vector<Regexp*> out;
      for (size_t i = 0; i <= sub.size(); i++) {
        std::vector<int32_t> istr;
        uint32_t iflags = 0;
        if (i < sub.size()) {
          const std::vector<int32_t>* p; uint32_t f;
          leading_string(sub[i], &p, &f);
          istr = *p; iflags = f;
          if (iflags == strflags) {
            size_t same = 0;
            while (same < str.size() && same < istr.size() && str[same] == istr[same]) same++;
            if (same > 0) { str.resize(same); continue; }
          }
        }
        if (i == start) {
        } else if (i == start + 1) {
          out.push_back(sub[start]);
        } else {
          Regexp* prefix = new_regexp(OpLiteral);
          prefix->flags = strflags;
          prefix->rune = str;
          for (size_t j = start; j < i; j++) sub[j] = remove_leading_string(sub[j], str.size());
          std::vector<Regexp*> run(sub.begin() + start, sub.begin() + i);
          Regexp* suffix = collapse(run, OpAlternate);
          Regexp* re = new_regexp(OpConcat);
          re->sub = {prefix, suffix};
          out.push_back(re);
        }
        start = i; str = istr; strflags = iflags;
      }
      sub = out;
    }

    // Round 2: Factor out common simple prefixes (first piece of each concatenation).
    {
      size_t start = 0;
      std::vector<Regexp*> out;
      Regexp* first = nullptr;
      for (size_t i = 0; i <= sub.size(); i++) {
        Regexp* ifirst = nullptr;
        if (i < sub.size()) {
          ifirst = leading_regexp(sub[i]);
          if (first != nullptr && regexp_equal(first, ifirst) &&
              (is_char_class(first) ||
               (first->op == OpRepeat && first->min == first->max && is_char_class(first->sub[0])))) {
            continue;
          }
        }
        if (i == start) {
        } else if (i == start + 1) {
          out.push_back(sub[start]);
        } else {
          Regexp* prefix = first;
          for (size_t j = start; j < i; j++) sub[j] = remove_leading_regexp(sub[j]);
          std::vector<Regexp*> run(sub.begin() + start, sub.begin() + i);
          Regexp* suffix = collapse(run, OpAlternate);
          Regexp* re = new_regexp(OpConcat);
          re->sub = {prefix, suffix};
          out.push_back(re);
        }
        start = i; first = ifirst;
      }
      sub = out;
    }

    // Round 3: Collapse runs of single literals or character classes.
    {
      size_t start = 0;
      std::vector<Regexp*> out;
      for (size_t i = 0; i <= sub.size(); i++) {
        if (i < sub.size() && is_char_class(sub[i])) continue;
        if (i == start) {
        } else if (i == start + 1) {
          out.push_back(sub[start]);
        } else {
          size_t mx = start;
          for (size_t j = start + 1; j < i; j++) {
            if (sub[mx]->op < sub[j]->op ||
                (sub[mx]->op == sub[j]->op && sub[mx]->rune.size() < sub[j]->rune.size()))
              mx = j;
          }
          std::swap(sub[start], sub[mx]);
          for (size_t j = start + 1; j < i; j++) merge_char_class(sub[start], sub[j]);
          clean_alt(sub[start]);
          out.push_back(sub[start]);
        }
        if (i < sub.size()) out.push_back(sub[i]);
        start = i + 1;
      }
      sub = out;
    }

    // Round 4: Collapse runs of empty matches into a single empty match.
    {
      std::vector<Regexp*> out;
      for (size_t i = 0; i < sub.size(); i++) {
        if (i + 1 < sub.size() && sub[i]->op == OpEmptyMatch && sub[i + 1]->op == OpEmptyMatch) continue;
        out.push_back(sub[i]);
      }
      sub = out;
    }
    return sub;
  }

  // parse.go: (*parser).swapVerticalBar / parseVerticalBar / parseRightParen
  bool swap_vertical_bar() {
    size_t n = stack.size();
    if (n >= 3 && stack[n - 2]->op == opVerticalBar && is_char_class(stack[n - 1]) && is_char_class(stack[n - 3])) {
      Regexp* re1 = stack[n - 1];
      Regexp* re3 = stack[n - 3];
      if (re1->op > re3->op) { std::swap(re1, re3); stack[n - 3] = re3; }
      merge_char_class(re3, re1);
      stack.pop_back();
      return true;
    }
    if (n >= 2) {
      Regexp* re1 = stack[n - 1];
      Regexp* re2 = stack[n - 2];
      if (re2->op == opVerticalBar) {
        if (n >= 3) clean_alt(stack[n - 3]);
        stack[n - 2] = re1;
        stack[n - 1] = re2;
        return true;
      }
    }
    return false;
  }

  void parse_vertical_bar() {
    concat();
    if (!swap_vertical_bar()) op(opVerticalBar);
  }

  bool parse_right_paren() {
    concat();
    if (swap_vertical_bar()) stack.pop_back();
    alternate();
    size_t n = stack.size();
    if (n < 2) { err = "unexpected )"; return false; }
    Regexp* re1 = stack[n - 1];
    Regexp* re2 = stack[n - 2];
    stack.resize(n - 2);
    if (re2->op != opLeftParen) { err = "unexpected )"; return false; }
    flags = re2->flags;
    if (re2->cap == 0) {
      push(re1);
    } else {
      re2->op = OpCapture;
      re2->sub.assign(1, re1);
      push(re2);
    }
    return true;
  }

  static bool is_valid_capture_name(const std::string& name) {
    if (name.empty()) return false;
    for (unsigned char c : name)
      if (c != '_' && !((c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'))) return false;
    return true;
  }

  // parse.go: (*parser).parsePerlFlags
  bool parse_perl_flags(const std::string& s, size_t& pos) {
    size_t rem = s.size() - pos;
    bool starts_p = rem > 4 && s[pos + 2] == 'P' && s[pos + 3] == '<';
    bool starts_name = rem > 3 && s[pos + 2] == '<';
    if (starts_p || starts_name) {
      size_t expr_start = pos + (starts_name ? 3 : 4);
      size_t end = s.find('>', pos);
      if (end == std::string::npos) { err = "invalid named capture"; return false; }
      std::string name = s.substr(expr_start, end - expr_start);
      if (!is_valid_capture_name(name)) { err = "invalid named capture"; return false; }
      num_cap++;
      Regexp* re = op(opLeftParen);
      re->cap = num_cap;
      re->name = name;
      pos = end + 1;
      return true;
    }
    size_t t = pos + 2;
    uint32_t fl = flags;
    int sign = +1;
    bool saw_flag = false;
    while (t < s.size()) {
      int32_t c;
      if (!next_rune(s, t, c, err)) return false;
      switch (c) {
        default: err = "invalid or unsupported Perl syntax"; return false;
        case 'i':
          err = "unsupported: case folding (?i)"; return false;
        case 'm': fl &= ~(uint32_t)OneLine; saw_flag = true; break;
        case 's': fl |= DotNL; saw_flag = true; break;
        case 'U': fl |= NonGreedy; saw_flag = true; break;
        case '-':
          if (sign < 0) { err = "invalid or unsupported Perl syntax"; return false; }
          sign = -1;
          fl = ~fl;
          saw_flag = false;
          break;
        case ':':
        case ')':
          if (sign < 0) {
            if (!saw_flag) { err = "invalid or unsupported Perl syntax"; return false; }
            fl = ~fl;
          }
          if (c == ':') op(opLeftParen);
          flags = fl;
          pos = t;
          return true;
      }
    }
    err = "missing closing )";
    return false;
  }

  static bool is_alnum(int32_t c) {
    return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z');
  }
  static int unhex(int32_t c) {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
  }

  // parse.go: (*parser).parseEscape.  pos points at the backslash.
  bool parse_escape(const std::string& s, size_t& pos, int32_t& out) {
    size_t t = pos + 1;
    if (t >= s.size()) { err = "trailing backslash at end of expression"; return false; }
    int32_t c;
    if (!next_rune(s, t, c, err)) return false;
    switch (c) {
      default:
        if (c < 0x80 && !is_alnum(c)) { out = c; pos = t; return true; }
        break;
      case '1': case '2': case '3': case '4': case '5': case '6': case '7':
        if (t >= s.size() || s[t] < '0' || s[t] > '7') break;
        /* fallthrough */
      case '0': {
        int32_t r = c - '0';
        for (int i = 1; i < 3; i++) {
          if (t >= s.size() || s[t] < '0' || s[t] > '7') break;
          r = r * 8 + (s[t] - '0');
          t++;
        }
        out = r; pos = t; return true;
      }
      case 'x': {
        if (t >= s.size()) break;
        if (!next_rune(s, t, c, err)) return false;
        if (c == '{') {
          int nhex = 0; int32_t r = 0; bool ok = false;
          while (true) {
            if (t >= s.size()) break;
            if (!next_rune(s, t, c, err)) return false;
            if (c == '}') { ok = true; break; }
            int v = unhex(c);
            if (v < 0) break;
            r = r * 16 + v;
            if (r > MaxRune) break;
            nhex++;
          }
          if (!ok || nhex == 0) break;
          out = r; pos = t; return true;
        }
        int x = unhex(c);
        if (t >= s.size()) break;
        if (!next_rune(s, t, c, err)) return false;
        int y = unhex(c);
        if (x < 0 || y < 0) break;
        out = x * 16 + y; pos = t; return true;
      }
      case 'a': out = 7; pos = t; return true;
      case 'f': out = '\f'; pos = t; return true;
      case 'n': out = '\n'; pos = t; return true;
      case 'r': out = '\r'; pos = t; return true;
      case 't': out = '\t'; pos = t; return true;
      case 'v': out = '\v'; pos = t; return true;
    }
    err = "invalid escape sequence";
    return false;
  }

  bool parse_class_char(const std::string& s, size_t& pos, int32_t& out) {
    if (pos >= s.size()) { err = "missing closing ]"; return false; }
    if (s[pos] == '\\') return parse_escape(s, pos, out);
    return next_rune(s, pos, out, err);
  }

  // parse.go: (*parser).parsePerlClassEscape.  Returns true if it consumed \d \s \w \D \S \W.
  bool parse_perl_class_escape(const std::string& s, size_t& pos, std::vector<int32_t>& cls) {
    if (!(flags & PerlX) || s.size() - pos < 2 || s[pos] != '\\') return false;
    const std::vector<int32_t>* g = nullptr; int sign = +1;
    switch (s[pos + 1]) {
      case 'd': g = &kPerlD; break;
      case 'D': g = &kPerlD; sign = -1; break;
      case 's': g = &kPerlS; break;
      case 'S': g = &kPerlS; sign = -1; break;
      case 'w': g = &kPerlW; break;
      case 'W': g = &kPerlW; sign = -1; break;
      default: return false;
    }
    if (sign < 0) append_negated_class(cls, *g); else append_class(cls, *g);
    pos += 2;
    return true;
  }

  // parse.go: (*parser).parseUnicodeClass.  pos at '\\'; returns true if it consumed \pN, \p{Name}, \P.. (err set on a
  // bad name: once the escape is seen the parser is committed).  Folded classes ((?i)) never get here: FoldCase is rejected.
  bool parse_unicode_class(const std::string& s, size_t& pos, std::vector<int32_t>& cls) {
    if (!(flags & UnicodeGroups) || s.size() - pos < 2 || s[pos] != '\\' || (s[pos + 1] != 'p' && s[pos + 1] != 'P')) return false;
    int sign = s[pos + 1] == 'P' ? -1 : +1;
    size_t t = pos + 2;
    if (t >= s.size()) { err = "invalid character class range"; return true; }
    std::string name;
    if (s[t] != '{') {
      // single-letter name (one rune)
      int32_t c;
      size_t t2 = t;
      if (!next_rune(s, t2, c, err)) return true;
      name = s.substr(t, t2 - t);
      t = t2;
    } else {
      const size_t end = s.find('}', pos);
      if (end == std::string::npos) { err = "invalid character class range"; return true; }
      name = s.substr(pos + 3, end - (pos + 3));
      t = end + 1;
    }
    if (!name.empty() && name[0] == '^') { sign = -sign; name = name.substr(1); }
    std::vector<int32_t> tab;
    if (!unicode_table(name, tab)) { err = "invalid character class range"; return true; }
    if (sign > 0) append_class(cls, tab); else append_negated_class(cls, tab);
    pos = t;
    return true;
  }

  // parse.go: (*parser).parseClass.  pos points at '['.
  bool parse_class(const std::string& s, size_t& pos) {
    size_t t = pos + 1;
    Regexp* re = new_regexp(OpCharClass);
    re->flags = flags;
    int sign = +1;
    if (t < s.size() && s[t] == '^') {
      sign = -1;
      t++;
      if (!(flags & ClassNL)) { re->rune.push_back('\n'); re->rune.push_back('\n'); }
    }
    std::vector<int32_t>& cls = re->rune;
    bool first = true;
    while (t >= s.size() || s[t] != ']' || first) {
      if (t >= s.size()) { err = "missing closing ]"; return false; }
      if (s[t] == '-' && !(flags & PerlX) && !first && (t + 1 == s.size() || s[t + 1] != ']')) {
        err = "invalid character class range"; return false;
      }
      first = false;
      if (s.size() - t > 2 && s[t] == '[' && s[t + 1] == ':') {
        size_t close = s.find(":]", t + 2);
        if (close != std::string::npos) {
          std::string name = s.substr(t, close + 2 - t);
          int sgn = +1;
          std::string key = name;
          if (key.size() > 3 && key[2] == '^') { sgn = -1; key = "[:" + key.substr(3); }
          const PosixGroup* found = nullptr;
          for (const auto& g : posix_groups()) if (key == g.name) found = &g;
          if (!found) { err = "invalid character class range"; return false; }
          if (sgn < 0) append_negated_class(cls, found->cls); else append_class(cls, found->cls);
          t = close + 2;
          continue;
        }
      }
      if (parse_unicode_class(s, t, cls)) { if (!err.empty()) return false; continue; }
      if (parse_perl_class_escape(s, t, cls)) continue;
      int32_t lo, hi;
      if (!parse_class_char(s, t, lo)) return false;
      hi = lo;
      if (s.size() - t >= 2 && s[t] == '-' && s[t + 1] != ']') {
        t++;
        if (!parse_class_char(s, t, hi)) return false;
        if (hi < lo) { err = "invalid character class range"; return false; }
      }
      append_range(cls, lo, hi);
    }
    t++;  // chop ]
    clean_class(cls);
    if (sign < 0) negate_class(cls);
    push(re);
    pos = t;
    return true;
  }

  // parse.go: (*parser).parseRepeat / parseInt.  pos at '{'.
  static bool parse_int(const std::string& s, size_t& t, int& n) {
    if (t >= s.size() || s[t] < '0' || s[t] > '9') return false;
    if (s.size() - t >= 2 && s[t] == '0' && s[t + 1] >= '0' && s[t + 1] <= '9') return false;
    size_t t0 = t;
    while (t < s.size() && s[t] >= '0' && s[t] <= '9') t++;
    n = 0;
    for (size_t i = t0; i < t; i++) {
      if (n >= 100000000) { n = -1; break; }
      n = n * 10 + (s[i] - '0');
    }
    return true;
  }

  static bool parse_repeat(const std::string& s, size_t pos, int& mn, int& mx, size_t& rest) {
    size_t t = pos;
    if (t >= s.size() || s[t] != '{') return false;
    t++;
    if (!parse_int(s, t, mn)) return false;
    if (t >= s.size()) return false;
    if (s[t] != ',') {
      mx = mn;
    } else {
      t++;
      if (t >= s.size()) return false;
      if (s[t] == '}') {
        mx = -1;
      } else {
        if (!parse_int(s, t, mx)) return false;
        if (mx < 0) mn = -1;
      }
    }
    if (t >= s.size() || s[t] != '}') return false;
    rest = t + 1;
    return true;
  }

  // parse.go: parse()
  Regexp* run(const std::string& s) {
    size_t t = 0;
    bool last_repeat = false;
    while (t < s.size()) {
      bool repeat_op = false;
      switch (s[t]) {
        default: {
          int32_t c;
          if (!next_rune(s, t, c, err)) return nullptr;
          literal(c);
          break;
        }
        case '(':
          if ((flags & PerlX) && s.size() - t >= 2 && s[t + 1] == '?') {
            if (!parse_perl_flags(s, t)) return nullptr;
            break;
          }
          num_cap++;
          op(opLeftParen)->cap = num_cap;
          t++;
          break;
        case '|':
          parse_vertical_bar();
          t++;
          break;
        case ')':
          if (!parse_right_paren()) return nullptr;
          t++;
          break;
        case '^':
          if (flags & OneLine) op(OpBeginText); else op(OpBeginLine);
          t++;
          break;
        case '$':
          if (flags & OneLine) op(OpEndText)->flags |= WasDollar; else op(OpEndLine);
          t++;
          break;
        case '.':
          if (flags & DotNL) op(OpAnyChar); else op(OpAnyCharNotNL);
          t++;
          break;
        case '[':
          if (!parse_class(s, t)) return nullptr;
          break;
        case '*':
        case '+':
        case '?': {
          int o = s[t] == '*' ? OpStar : (s[t] == '+' ? OpPlus : OpQuest);
          t++;
          if (!repeat(o, 0, 0, s, t, last_repeat)) return nullptr;
          repeat_op = true;
          break;
        }
        case '{': {
          int mn, mx; size_t rest;
          if (!parse_repeat(s, t, mn, mx, rest)) {
            literal('{');
            t++;
            break;
          }
          if (mn < 0 || mn > 1000 || mx > 1000 || (mx >= 0 && mn > mx)) { err = "invalid repeat count"; return nullptr; }
          t = rest;
          if (!repeat(OpRepeat, mn, mx, s, t, last_repeat)) return nullptr;
          repeat_op = true;
          break;
        }
        case '\\': {
          bool handled = false;
          if ((flags & PerlX) && s.size() - t >= 2) {
            switch (s[t + 1]) {
              case 'A': op(OpBeginText); t += 2; handled = true; break;
              case 'b': op(OpWordBoundary); t += 2; handled = true; break;
              case 'B': op(OpNoWordBoundary); t += 2; handled = true; break;
              case 'C': err = "invalid escape sequence \\C"; return nullptr;
              case 'Q': {
                size_t e = s.find("\\E", t + 2);
                std::string lit = s.substr(t + 2, e == std::string::npos ? std::string::npos : e - (t + 2));
                size_t lp = 0;
                while (lp < lit.size()) {
                  int32_t c;
                  if (!next_rune(lit, lp, c, err)) return nullptr;
                  literal(c);
                }
                t = (e == std::string::npos) ? s.size() : e + 2;
                handled = true;
                break;
              }
              case 'z': op(OpEndText); t += 2; handled = true; break;
            }
          }
          if (handled) break;
          Regexp* re = new_regexp(OpCharClass);
          re->flags = flags;
          if (s.size() - t >= 2 && (s[t + 1] == 'p' || s[t + 1] == 'P')) {
            // parse.go: the escape site of parseUnicodeClass
            if (parse_unicode_class(s, t, re->rune)) {
              if (!err.empty()) return nullptr;
              push(re);
              break;
            }
          }
          if (parse_perl_class_escape(s, t, re->rune)) { push(re); break; }
          int32_t c;
          if (!parse_escape(s, t, c)) return nullptr;
          literal(c);
          break;
        }
      }
      if (!err.empty()) return nullptr;
      last_repeat = repeat_op;
    }
    concat();
    if (swap_vertical_bar()) stack.pop_back();
    alternate();
    if (!err.empty()) return nullptr;
    if (stack.size() != 1) { err = "missing closing )"; return nullptr; }
    return stack[0];
  }
};

bool parse(const std::string& pattern, uint32_t flags, Arena& arena, Regexp** out, std::string& err) {
  Parser p(arena, flags);
  Regexp* re = p.run(pattern);
  if (re == nullptr || !p.err.empty()) { err = p.err.empty() ? "parse error" : p.err; return false; }
  *out = re;
  return true;
}

// ---------------------------------------------------------------------------------------
// simplify.go
static Regexp* simplify1(int op, uint32_t flags, Regexp* sub, Regexp* re, Arena& arena) {
  if (sub->op == OpEmptyMatch) return sub;
  if (op == sub->op && (flags & NonGreedy) == (sub->flags & NonGreedy)) return sub;
  if (re != nullptr && re->op == op && (re->flags & NonGreedy) == (flags & NonGreedy) && sub == re->sub[0]) return re;
  Regexp* nre = arena.make(op);
  nre->flags = flags;
  nre->sub.assign(1, sub);
  return nre;
}

Regexp* simplify(Regexp* re, Arena& arena) {
  if (re == nullptr) return nullptr;
  switch (re->op) {
    case OpCapture:
    case OpConcat:
    case OpAlternate: {
      Regexp* nre = re;
      for (size_t i = 0; i < re->sub.size(); i++) {
        Regexp* sub = re->sub[i];
        Regexp* nsub = simplify(sub, arena);
        if (nre == re && nsub != sub) {
          nre = arena.make(re->op);
          nre->flags = re->flags; nre->min = re->min; nre->max = re->max; nre->cap = re->cap; nre->name = re->name;
          nre->sub.assign(re->sub.begin(), re->sub.begin() + i);
        }
        if (nre != re) nre->sub.push_back(nsub);
      }
      return nre;
    }
    case OpStar:
    case OpPlus:
    case OpQuest: {
      Regexp* sub = simplify(re->sub[0], arena);
      return simplify1(re->op, re->flags, sub, re, arena);
    }
    case OpRepeat: {
      if (re->min == 0 && re->max == 0) return arena.make(OpEmptyMatch);
      Regexp* sub = simplify(re->sub[0], arena);
      if (re->max == -1) {
        if (re->min == 0) return simplify1(OpStar, re->flags, sub, nullptr, arena);
        if (re->min == 1) return simplify1(OpPlus, re->flags, sub, nullptr, arena);
        Regexp* nre = arena.make(OpConcat);
        for (int i = 0; i < re->min - 1; i++) nre->sub.push_back(sub);
        nre->sub.push_back(simplify1(OpPlus, re->flags, sub, nullptr, arena));
        return nre;
      }
      if (re->min == 1 && re->max == 1) return sub;
      Regexp* prefix = nullptr;
      if (re->min > 0) {
        prefix = arena.make(OpConcat);
        for (int i = 0; i < re->min; i++) prefix->sub.push_back(sub);
      }
      if (re->max > re->min) {
        Regexp* suffix = simplify1(OpQuest, re->flags, sub, nullptr, arena);
        for (int i = re->min + 1; i < re->max; i++) {
          Regexp* nre2 = arena.make(OpConcat);
          nre2->sub = {sub, suffix};
          suffix = simplify1(OpQuest, re->flags, nre2, nullptr, arena);
        }
        if (prefix == nullptr) return suffix;
        prefix->sub.push_back(suffix);
      }
      if (prefix != nullptr) return prefix;
      return arena.make(OpNoMatch);
    }
  }
  return re;
}

// ---------------------------------------------------------------------------------------
// compile.go
namespace {

struct PatchList { uint32_t head = 0, tail = 0; };
struct Frag { uint32_t i = 0; PatchList out; bool nullable = false; };

struct Compiler {
  Prog& p;
  explicit Compiler(Prog& prog) : p(prog) {}

  static PatchList make_patch_list(uint32_t n) { return PatchList{n, n}; }

  void patch(PatchList l, uint32_t val) {
    uint32_t head = l.head;
    while (head != 0) {
      Inst& i = p.inst[head >> 1];
      if ((head & 1) == 0) { head = i.out; i.out = val; }
      else { head = i.arg; i.arg = val; }
    }
  }

  PatchList append(PatchList l1, PatchList l2) {
    if (l1.head == 0) return l2;
    if (l2.head == 0) return l1;
    Inst& i = p.inst[l1.tail >> 1];
    if ((l1.tail & 1) == 0) i.out = l2.head; else i.arg = l2.head;
    return PatchList{l1.head, l2.tail};
  }

  Frag inst(uint8_t op) {
    Frag f; f.i = (uint32_t)p.inst.size(); f.nullable = true;
    Inst in; in.op = op;
    p.inst.push_back(in);
    return f;
  }
  Frag nop() { Frag f = inst(InstNop); f.out = make_patch_list(f.i << 1); return f; }
  Frag fail() { return Frag(); }
  Frag cap(uint32_t arg) {
    Frag f = inst(InstCapture);
    f.out = make_patch_list(f.i << 1);
    p.inst[f.i].arg = arg;
    if (p.num_cap < (int)arg + 1) p.num_cap = (int)arg + 1;
    return f;
  }
  Frag cat(Frag f1, Frag f2) {
    if (f1.i == 0 || f2.i == 0) return Frag();
    patch(f1.out, f2.i);
    Frag f; f.i = f1.i; f.out = f2.out; f.nullable = f1.nullable && f2.nullable;
    return f;
  }
  Frag alt(Frag f1, Frag f2) {
    if (f1.i == 0) return f2;
    if (f2.i == 0) return f1;
    Frag f = inst(InstAlt);
    p.inst[f.i].out = f1.i;
    p.inst[f.i].arg = f2.i;
    f.out = append(f1.out, f2.out);
    f.nullable = f1.nullable || f2.nullable;
    return f;
  }
  Frag quest(Frag f1, bool nongreedy) {
    Frag f = inst(InstAlt);
    if (nongreedy) { p.inst[f.i].arg = f1.i; f.out = make_patch_list(f.i << 1); }
    else { p.inst[f.i].out = f1.i; f.out = make_patch_list(f.i << 1 | 1); }
    f.out = append(f.out, f1.out);
    return f;
  }
  Frag loop(Frag f1, bool nongreedy) {
    Frag f = inst(InstAlt);
    if (nongreedy) { p.inst[f.i].arg = f1.i; f.out = make_patch_list(f.i << 1); }
    else { p.inst[f.i].out = f1.i; f.out = make_patch_list(f.i << 1 | 1); }
    patch(f1.out, f.i);
    return f;
  }
  Frag star(Frag f1, bool nongreedy) {
    if (f1.nullable) return quest(plus(f1, nongreedy), nongreedy);
    return loop(f1, nongreedy);
  }
  Frag plus(Frag f1, bool nongreedy) {
    Frag f; f.i = f1.i; f.out = loop(f1, nongreedy).out; f.nullable = f1.nullable;
    return f;
  }
  Frag empty(uint32_t op) {
    Frag f = inst(InstEmptyWidth);
    p.inst[f.i].arg = op;
    f.out = make_patch_list(f.i << 1);
    return f;
  }
  Frag rune(const std::vector<int32_t>& r, uint32_t flags) {
    Frag f = inst(InstRune);
    f.nullable = false;
    Inst& i = p.inst[f.i];
    i.rune = r;
    flags &= FoldCase;
    if (r.size() != 1 || simple_fold(r[0]) == r[0]) flags &= ~(uint32_t)FoldCase;
    i.arg = flags;
    f.out = make_patch_list(f.i << 1);
    if ((flags & FoldCase) == 0 && (r.size() == 1 || (r.size() == 2 && r[0] == r[1]))) i.op = InstRune1;
    else if (r.size() == 2 && r[0] == 0 && r[1] == MaxRune) i.op = InstRuneAny;
    else if (r.size() == 4 && r[0] == 0 && r[1] == '\n' - 1 && r[2] == '\n' + 1 && r[3] == MaxRune) i.op = InstRuneAnyNotNL;
    return f;
  }

  Frag compile(Regexp* re) {
    static const std::vector<int32_t> any_rune_not_nl = {0, '\n' - 1, '\n' + 1, MaxRune};
    static const std::vector<int32_t> any_rune = {0, MaxRune};
    switch (re->op) {
      case OpNoMatch: return fail();
      case OpEmptyMatch: return nop();
      case OpLiteral: {
        if (re->rune.empty()) return nop();
        Frag f;
        for (size_t j = 0; j < re->rune.size(); j++) {
          std::vector<int32_t> one(1, re->rune[j]);
          Frag f1 = rune(one, re->flags);
          if (j == 0) f = f1; else f = cat(f, f1);
        }
        return f;
      }
      case OpCharClass: return rune(re->rune, re->flags);
      case OpAnyCharNotNL: return rune(any_rune_not_nl, 0);
      case OpAnyChar: return rune(any_rune, 0);
      case OpBeginLine: return empty(EmptyBeginLine);
      case OpEndLine: return empty(EmptyEndLine);
      case OpBeginText: return empty(EmptyBeginText);
      case OpEndText: return empty(EmptyEndText);
      case OpWordBoundary: return empty(EmptyWordBoundary);
      case OpNoWordBoundary: return empty(EmptyNoWordBoundary);
      case OpCapture: {
        Frag bra = cap((uint32_t)(re->cap << 1));
        Frag sub = compile(re->sub[0]);
        Frag ket = cap((uint32_t)(re->cap << 1 | 1));
        return cat(cat(bra, sub), ket);
      }
      case OpStar: return star(compile(re->sub[0]), (re->flags & NonGreedy) != 0);
      case OpPlus: return plus(compile(re->sub[0]), (re->flags & NonGreedy) != 0);
      case OpQuest: return quest(compile(re->sub[0]), (re->flags & NonGreedy) != 0);
      case OpConcat: {
        if (re->sub.empty()) return nop();
        Frag f;
        for (size_t i = 0; i < re->sub.size(); i++) {
          if (i == 0) f = compile(re->sub[i]); else f = cat(f, compile(re->sub[i]));
        }
        return f;
      }
      case OpAlternate: {
        Frag f;
        for (Regexp* sub : re->sub) f = alt(f, compile(sub));
        return f;
      }
    }
    return Frag();
  }
};

}  // namespace

bool compile(Regexp* re, Prog& prog, std::string& err) {
  prog.inst.clear();
  prog.num_cap = 2;
  Compiler c(prog);
  c.inst(InstFail);
  Frag f = c.compile(re);
  Frag m = c.inst(InstMatch);
  c.patch(f.out, m.i);
  prog.start = (int)f.i;
  for (const Inst& in : prog.inst) {
    if ((in.op == InstRune || in.op == InstRune1) && (in.arg & FoldCase)) {
      err = "unsupported: case-folded rune instruction";
      return false;
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------
static void dump(const Regexp* re, std::string& o) {
  static const char* names[] = {"?", "nomatch", "empty", "lit", "cc", "anynotnl", "any", "bol", "eol", "bot",
                                "eot", "wb", "nwb", "cap", "star", "plus", "quest", "repeat", "cat", "alt"};
  o += names[re->op < 20 ? re->op : 0];
  if (re->flags & NonGreedy) o += "?";
  if (re->op == OpLiteral || re->op == OpCharClass) {
    o += "[";
    for (size_t i = 0; i < re->rune.size(); i++) { if (i) o += ","; o += std::to_string(re->rune[i]); }
    o += "]";
  }
  if (re->op == OpCapture) o += "#" + std::to_string(re->cap) + (re->name.empty() ? "" : ":" + re->name);
  if (re->op == OpRepeat) o += "{" + std::to_string(re->min) + "," + std::to_string(re->max) + "}";
  if (!re->sub.empty()) {
    o += "(";
    for (size_t i = 0; i < re->sub.size(); i++) { if (i) o += " "; dump(re->sub[i], o); }
    o += ")";
  }
}

std::string regexp_to_string(const Regexp* re) { std::string o; dump(re, o); return o; }

}  // namespace rgx
