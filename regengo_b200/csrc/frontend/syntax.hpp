// Host-side front-end: restatement of Go's regexp/syntax (Parse(Perl) -> Simplify -> Compile).
//
// The reference calls the Go standard library here (regengo.go:92,98,104); that library is
// NOT vendored under /root/reference (go.mod:3 pins only "go 1.24").  This file restates the
// published algorithm of regexp/syntax {parse.go, simplify.go, compile.go} so that the
// instruction program (syntax.Prog) has the identical layout: instruction numbering, Alt
// Out/Arg priority, capture placement, literal merging, alternation factoring and the
// {n,m} expansion all feed the reference's engine selection, its TDFA tables and the
// failure offsets of its skip-restart quirk, so they must match exactly.
// The layout is pinned by tests/test_frontend_goldens.py against the programs embedded in the
// reference's checked-in generated files.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace rgx {

// regexp/syntax Op values (parse-time ordering is significant: swapVerticalBar and
// factor() round 3 compare Op numerically).
enum Op : int {
  OpNoMatch = 1, OpEmptyMatch, OpLiteral, OpCharClass, OpAnyCharNotNL, OpAnyChar,
  OpBeginLine, OpEndLine, OpBeginText, OpEndText, OpWordBoundary, OpNoWordBoundary,
  OpCapture, OpStar, OpPlus, OpQuest, OpRepeat, OpConcat, OpAlternate,
  opPseudo = 128, opLeftParen = 128, opVerticalBar = 129
};

enum Flags : uint32_t {
  FoldCase = 1, Literal = 2, ClassNL = 4, DotNL = 8, OneLine = 16, NonGreedy = 32,
  PerlX = 64, UnicodeGroups = 128, WasDollar = 256, Simple = 512,
  Perl = ClassNL | OneLine | PerlX | UnicodeGroups
};

struct Regexp {
  int op = 0;
  uint32_t flags = 0;
  std::vector<Regexp*> sub;
  std::vector<int32_t> rune;  // literal runes, or [lo,hi] pairs for a class
  int min = 0, max = 0;
  int cap = 0;
  std::string name;
};

// syntax.InstOp values.
enum InstOp : uint8_t {
  InstAlt = 0, InstAltMatch, InstCapture, InstEmptyWidth, InstMatch, InstFail, InstNop,
  InstRune, InstRune1, InstRuneAny, InstRuneAnyNotNL
};

enum EmptyOp : uint32_t {
  EmptyBeginLine = 1, EmptyEndLine = 2, EmptyBeginText = 4, EmptyEndText = 8,
  EmptyWordBoundary = 16, EmptyNoWordBoundary = 32
};

struct Inst {
  uint8_t op = InstFail;
  uint32_t out = 0, arg = 0;
  std::vector<int32_t> rune;
};

struct Prog {
  std::vector<Inst> inst;
  int start = 0;
  int num_cap = 2;
};

// Owns every Regexp node created while parsing / simplifying one pattern.
struct Arena {
  std::vector<Regexp*> nodes;
  ~Arena() { for (auto* n : nodes) delete n; }
  Regexp* make(int op) { auto* r = new Regexp(); r->op = op; nodes.push_back(r); return r; }
};

// Each returns false and fills err on failure.
bool parse(const std::string& pattern, uint32_t flags, Arena& arena, Regexp** out, std::string& err);
Regexp* simplify(Regexp* re, Arena& arena);
bool compile(Regexp* re, Prog& prog, std::string& err);

// unicode.IsLetter / unicode.IsDigit over the same tables as \p{L} / \p{Nd} (used by the replace-template parser)
bool rune_is_letter(int32_t r);
bool rune_is_digit(int32_t r);

std::string regexp_to_string(const Regexp* re);  // debug dump (not Go's String())

}  // namespace rgx
