// The compiled "program" for one pattern: the syntax.Prog plus every decision the reference's
// generator makes from it (engine per method, memoisation, checkpoint mode, prefix byte,
// Thompson masks, TDFA tables, stream defaults).  This is what the device kernels execute and
// what a retargeted Go generator would emit as a blob.
//
// Restates: regengo.go:86-156 (pipeline), internal/compiler/compiler.go:59-184 (New +
// analyzeAndLog), analysis.go (all predicates), thompson.go:25-66, tdfa.go:83-290 (buildTDFA),
// analysis_match_len.go, streaming.go:25-62 (stream defaults), analyze_api.go:137-218 (labels).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "syntax.hpp"

namespace rgx {

struct TagAction { int tag; int offset; };
inline bool operator==(const TagAction& a, const TagAction& b) { return a.tag == b.tag && a.offset == b.offset; }
inline bool operator!=(const TagAction& a, const TagAction& b) { return !(a == b); }

struct Tdfa {
  bool built = false;
  int num_states = 0;
  int num_tags = 0;       // getTagCount(): 2 * len(captureNames)
  int start_begin = 0, start_any = 0;
  std::vector<int> init_tags_begin, init_tags_any;     // tag ids set to `start`
  std::vector<int32_t> trans;                          // [ns*128], -1 = none
  std::vector<std::vector<TagAction>> actions;         // [ns*128]
  std::vector<uint8_t> accept, accept_eot;             // [ns]
  std::vector<std::vector<TagAction>> accept_actions;  // [ns]
  int max_actions = 0, max_accept_actions = 0;
};

enum MatchEngine { MATCH_BT = 0, MATCH_THOMPSON = 1 };
enum FindEngine { FIND_NONE = 0, FIND_BT = 1, FIND_TDFA = 2 };

struct Options {
  bool force_thompson = false, force_tnfa = false, force_tdfa = false;
  int tdfa_threshold = 0;  // 0 => 500
};

struct Program {
  std::string pattern;
  Options opts;
  Prog prog;
  std::vector<std::string> capture_names;  // index 0 = ""
  // analysis.go / compiler.go:59-90
  bool has_captures = false, needs_backtracking = false, anchored = false, has_word_boundary = false;
  bool nested_loops = false, catastrophic_risk = false, end_anchor = false, use_thompson_nfa = false;
  // engine selection (compiler.go:93-184, :262-325, :423-426)
  int match_engine = MATCH_BT;
  bool match_memo = false;       // useMemoization while Match* is generated
  int find_engine = FIND_NONE;
  bool find_memo = false;        // useMemoization while Find*/FindAll* are generated
  bool tnfa = false;             // "TNFA" = memoised backtracking for captures
  bool per_capture_ckpt = false;
  std::vector<uint8_t> alt_ckpt;       // [n_inst] altsNeedingCheckpoint
  std::vector<uint8_t> greedy_loop;    // [n_inst] Match-mode "simple greedy loop" Alts (instructions.go:411-436)
  bool has_prefix = false; uint8_t prefix = 0;  // findRequiredPrefix (compiler.go:719-737)
  // per-instruction byte classes: 256-bit membership for Rune insts whose ranges are all < 128
  // (charclass.go:10-18, 43-54); unicode_class[i] = 1 when the class has a range bound >= 128.
  std::vector<uint8_t> unicode_class;  // [n_inst]
  std::vector<uint32_t> class_bits;    // [n_inst*8]; for unicode classes: ASCII part only
  // analysis_match_len.go + streaming.go
  int min_match_len = 0, max_match_len = 0;
  int default_max_leftover = 0, min_buffer = 0;
  // Thompson (thompson.go:25-60, analysis.go:447-501)
  uint64_t start_closure = 0, accept_mask = 0;
  std::vector<uint64_t> closures;      // [n_inst]
  std::vector<uint64_t> eps_after;     // [n_inst] = closures[Inst[i].Out] for char states else 0
  std::vector<uint8_t> char_state;     // [n_inst]
  std::vector<uint32_t> thompson_cond; // [n_inst*8] byte set of the per-state condition (thompson.go:199-303)
  // TDFA
  Tdfa tdfa;
  // Analyze() labels (analyze_api.go)
  std::vector<std::string> engine_labels, feature_labels;
};

// Full pipeline; returns false + err for unsupported/invalid patterns.
bool build_program(const std::string& pattern, const Options& opts, Program& out, std::string& err);

std::string program_to_json(const Program& p);

}  // namespace rgx
