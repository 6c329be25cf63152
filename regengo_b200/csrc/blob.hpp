// The packed per-pattern "program blob": a flat array of little-endian u32 words holding the
// syntax.Prog plus every generator decision (engine per method, memoisation, checkpoint mode,
// prefix byte, Thompson masks, TDFA tables, stream defaults).  It is what a retargeted
// internal/compiler would emit in place of the Go goto-machine and what
// rgx_load() consumes.  The oracle under oracle/ carries its own independent reader of this
// layout (oracle/rgx_oracle.c); tests/golden/golden_blob.py writes it straight from the programs
// mined out of the reference's generated Go files.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "frontend/program.hpp"

namespace rgx {

constexpr uint32_t BLOB_MAGIC = 0x42584752u;  // "RGXB"
constexpr uint32_t BLOB_VERSION = 1;

// header word indices
enum BlobHdr : int {
  H_MAGIC = 0, H_VERSION, H_WORDS, H_NINST, H_START, H_NUMCAP, H_FLAGS, H_PREFIX,
  H_MATCH_ENGINE, H_FIND_ENGINE, H_MINLEN, H_MAXLEN, H_LEFTOVER, H_MINBUF,
  H_OFF_INST,      // n_inst * 4 words: {op | iflags<<8, out, arg, aux}
  H_OFF_CLASS,     // n_inst * 8 words: 256-bit byte set of a Rune inst (ASCII part for unicode classes)
  H_OFF_RANGES,    // n_inst * 2 words {first pair index, pair count} then the [lo,hi] pairs
  H_NRANGE_PAIRS,
  H_OFF_THOMPSON,  // start_closure(2) accept_mask(2) eps_after[n_inst](2 each) cond[n_inst](8 each)
  H_OFF_TDFA,      // 0 when absent; see blob.cpp
  H_OFF_NAMES,     // n_groups+1 length-prefixed names, byte-packed
  H_NAMES_WORDS,
  H_WORDS_HDR = 32
};

// H_FLAGS bits
enum BlobFlags : uint32_t {
  F_ANCHORED = 1u << 0, F_NEEDS_BT = 1u << 1, F_HAS_PREFIX = 1u << 2, F_MATCH_MEMO = 1u << 3,
  F_FIND_MEMO = 1u << 4, F_PER_CAPTURE = 1u << 5, F_HAS_CAPTURES = 1u << 6, F_WORD_BOUNDARY = 1u << 7,
};

// per-instruction flag bits (byte 1 of inst word 0)
enum InstFlags : uint32_t { IF_ALT_CKPT = 1, IF_GREEDY_LOOP = 2, IF_UNICODE_CLASS = 4, IF_CHAR_STATE = 8,
                            IF_ATOMIC_LOOP = 16 /* device image only, see device_program.cu */ };

// TDFA section header (word offsets relative to H_OFF_TDFA)
enum TdfaHdr : int {
  T_NSTATES = 0, T_NTAGS, T_START_BEGIN, T_START_ANY, T_NINIT_BEGIN, T_NINIT_ANY,
  T_NACTS, T_NACC_ACTS, T_MAX_ACTS, T_MAX_ACC_ACTS, T_WORDS_HDR = 12
  // then: init_begin[n], init_any[n], trans[ns*128] (i32, -1 none), accept[ns], accept_eot[ns],
  //       act_off[ns*128+1], acts[nacts*2] {tag, offset}, acc_off[ns+1], acc_acts[nacc*2]
};

std::vector<uint32_t> program_to_blob(const Program& p);
bool blob_to_program(const uint32_t* w, size_t n_words, Program& out, std::string& err);

}  // namespace rgx
