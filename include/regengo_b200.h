/*
 * regengo_b200 -- C ABI of the B200 matching path.
 *
 * The reference (KromDaniel/regengo) has no FFI: its hot path is the method set that
 * `regengo.Compile` GENERATES per pattern (internal/compiler/compiler.go:204-367).  A drop-in
 * replaces the *bodies* of those generated methods with calls into this library (INTEGRATION.md
 * shows the cgo stub).  Every entry point below names the generated method it stands in for.
 *
 * Conventions
 *   - plain pointers and sizes only; the library never owns caller memory;
 *   - every function returns >= 0 on success and a negative RGX_E* code on failure; a message
 *     can be fetched with rgx_last_error() (thread-local);
 *   - there is NO CPU matching path: without a usable CUDA device every compute entry point
 *     returns RGX_ECUDA.  Only the compile/inspect functions work without a GPU;
 *   - a `rgx_program` is immutable after creation and may be shared between threads; a `rgx_ctx`
 *     owns one CUDA stream plus scratch and must not be used from two threads at once
 *     (mirrors the reference's stateless receivers + sync.Pool scratch, pool.go:11-169).
 *
 * Result layout ("offset records")
 *   One match = 2*(k+1) int64 byte offsets into the caller's input, group 0 first
 *   (k = number of capture groups): [start0,end0,start1,end1,...].  A group the reference
 *   would leave nil is (-1,-1).  This is the flat form of the generated `<Name>BytesResult`
 *   struct of []byte slices (captures.go:83-118): field i == input[start_i:end_i].
 */
#ifndef REGENGO_B200_H
#define REGENGO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RGX_OK 0
#define RGX_EINVAL (-1)      /* bad argument */
#define RGX_EPATTERN (-2)    /* pattern does not parse / unsupported construct */
#define RGX_EUNSUPPORTED (-3)/* method not generated for this pattern (e.g. Find* without captures) */
#define RGX_ECUDA (-4)       /* no device / CUDA failure */
#define RGX_ENOMEM (-5)
#define RGX_EBUFFER_TOO_SMALL (-6) /* stream.ErrBufferTooSmall (stream/stream.go:96-101) */
#define RGX_ECAPACITY (-7)   /* caller output buffer too small; see the function's doc */

typedef struct rgx_program rgx_program;
typedef struct rgx_ctx rgx_ctx;

/* regengo.Options fields that change matching semantics (regengo.go:13-65). */
typedef struct rgx_options {
  int32_t force_thompson;  /* Options.ForceThompson */
  int32_t force_tnfa;      /* Options.ForceTNFA     */
  int32_t force_tdfa;      /* Options.ForceTDFA     */
  int32_t tdfa_threshold;  /* Options.TDFAThreshold, 0 => 500 */
} rgx_options;

/* What the generator decided for a pattern (compiler.go:93-184). */
typedef struct rgx_info {
  int32_t n_inst;          /* len(prog.Inst) */
  int32_t num_cap;         /* prog.NumCap = 2*(k+1) */
  int32_t n_groups;        /* k */
  int32_t match_engine;    /* 0 backtracking, 1 Thompson bitset */
  int32_t find_engine;     /* 0 none (no captures => no Find* methods), 1 backtracking, 2 TDFA */
  int32_t match_memo;      /* visited bit-vector in Match*  */
  int32_t find_memo;       /* visited bit-vector in Find* ("TNFA") */
  int32_t per_capture_ckpt;/* per-capture (1) or array (0) checkpointing */
  int32_t anchored;
  int32_t min_match_len;   /* <Name>MinMatchLen */
  int32_t max_match_len;   /* <Name>MaxMatchLen, -1 unbounded */
  int32_t default_max_leftover; /* DefaultMaxLeftover() (streaming.go:25-46) */
  int32_t min_buffer;      /* cfg.Validate(minBuffer) (streaming.go:56-62) */
  int32_t tdfa_states;     /* 0 unless find_engine == 2 */
  int32_t tdfa_tags;
  int32_t reserved[5];
} rgx_info;

/* ---- generation time: stands in for regengo.Compile (regengo.go:86-156) ------------------ */

/* Parse(Perl) -> Simplify -> Compile -> analysis -> engine selection -> tables. Host only. */
int rgx_compile(const char* pattern, const rgx_options* opts /* may be NULL */, rgx_program** out);
void rgx_program_free(rgx_program* p);
int rgx_program_info(const rgx_program* p, rgx_info* out);
/* JSON dump of the program (Prog listing, masks, TDFA tables); the string lives as long as p. */
const char* rgx_program_json(const rgx_program* p);
/* JSON summary of how the device kernels will run the program (host computation, no GPU needed): image size,
 * FindAll start filter and prefix skip, run-anchor shape, straight-line forms, FindReader path.  Written into
 * buf (NUL-terminated) when cap suffices; returns the length needed (without the NUL). */
int64_t rgx_program_device_plan(const rgx_program* p, char* buf, size_t cap);
/* The packed device image (u32 words; the plan's w6_* entries are offsets into it).  Copies when cap_words suffices;
 * returns the number of words.  Host computation; used by the packing tests. */
int64_t rgx_program_device_image(const rgx_program* p, uint32_t* words_out, size_t cap_words);
/* Name of capture group i (1..k), "" if unnamed; NULL if out of range. */
const char* rgx_program_group_name(const rgx_program* p, int32_t i);

/* The packed per-pattern "program blob" a retargeted Go generator would embed in place of the
 * goto-machine.  rgx_program_blob: returns the blob size in bytes; copies it when cap suffices. */
int64_t rgx_program_blob(const rgx_program* p, void* buf, size_t cap);
int rgx_load(const void* blob, size_t n, rgx_program** out);

const char* rgx_last_error(void);
const char* rgx_version(void);

/* ---- device context ---------------------------------------------------------------------- */

int rgx_ctx_create(int32_t device, rgx_ctx** out);
void rgx_ctx_destroy(rgx_ctx* c);
/* Number of kernels this context has launched so far (bench.py's gpu_launches claim). */
int64_t rgx_ctx_launches(const rgx_ctx* c);
/* Statistics of the last call on this context.  which = 0: chunks of the last FindReader call that were replayed by
 * the sequential form of the chase instead of the lane-parallel one (-1: unknown `which`). */
int64_t rgx_ctx_stat(const rgx_ctx* c, int32_t which);
/* Upload granularity of the host-buffer FindAll (rgx_find_all / rgx_find_all_rle): inputs of at least two
 * chunks are uploaded chunk by chunk while earlier chunks are scanned (default 256 MiB; a multiple of
 * 32 KiB, at least 64 KiB). */
int rgx_ctx_set_chunk_bytes(rgx_ctx* c, uint64_t bytes);
/* Per-phase device timing of the last rgx_find_all*_dev call (CUDA events recorded on the context's
 * stream): out_ms[4] = {scan kernel, chain kernels, compaction kernels, all three}. */
int rgx_ctx_enable_timing(rgx_ctx* c, int32_t on);
int rgx_ctx_last_timing(const rgx_ctx* c, float* out_ms);
/* The context's CUDA stream as an opaque cudaStream_t, for event timing on the launching stream. */
void* rgx_ctx_stream(const rgx_ctx* c);
int rgx_ctx_sync(rgx_ctx* c);

/* ---- MatchBytes: func (T) MatchBytes(input []byte) bool  (compiler.go:304-307, 740-871;
 *      Thompson variant thompson.go:69-131).  Batch extension: input i = bytes[offs[i]:offs[i+1]],
 *      out[i] = 0/1.  Host pointers.                                                            */
int rgx_match_batch(rgx_ctx* c, const rgx_program* p, const uint8_t* bytes, const uint64_t* offs,
                    uint64_t n, uint8_t* out);
/* Same with DEVICE pointers (inputs already resident in HBM); asynchronous on the ctx stream. */
int rgx_match_batch_dev(rgx_ctx* c, const rgx_program* p, const uint8_t* d_bytes,
                        const uint64_t* d_offs, uint64_t n, uint8_t* d_out);

/* Many programs, ONE launch (BASELINE.json configs[3]: a pattern suite over a packed batch of short inputs).  The
 * inputs of program p are the index range [prog_first[p], prog_first[p+1]) of the batch (prog_first is a HOST array of
 * n_progs + 1 entries); out[i - prog_first[0]] = MatchBytes of progs[p] on input i.  Host form: 64-bit offsets as in
 * rgx_match_batch, uploaded as they are (the kernel has a 64-bit-offset form).
 * Device form: d_offs32[i] is relative to d_bytes, d_out is indexed by i. */
int rgx_match_multi(rgx_ctx* c, const rgx_program* const* progs, uint32_t n_progs, const uint8_t* bytes,
                    const uint64_t* offs, const uint64_t* prog_first, uint8_t* out);
int rgx_match_multi_dev(rgx_ctx* c, const rgx_program* const* progs, uint32_t n_progs, const uint8_t* d_bytes,
                        const uint32_t* d_offs32, const uint64_t* prog_first, uint8_t* d_out);

/* ---- FindBytes: func (T) FindBytes(input []byte) (*TBytesResult, bool)  (find.go:41-67,
 *      469-591; TDFA compiler.go:505-534 + tdfa.go:831-1052).  Batch extension: found[i] = 0/1,
 *      out_offsets[i*num_cap ..] = offset record relative to input i (undefined if !found).     */
int rgx_find_batch(rgx_ctx* c, const rgx_program* p, const uint8_t* bytes, const uint64_t* offs,
                   uint64_t n, uint8_t* found, int64_t* out_offsets);
int rgx_find_batch_dev(rgx_ctx* c, const rgx_program* p, const uint8_t* d_bytes,
                       const uint64_t* d_offs, uint64_t n, uint8_t* d_found, int64_t* d_out_offsets);

/* ---- FindAllBytes: func (T) FindAllBytes(input []byte, n int) []*TBytesResult
 *      (backtracking find.go:100-316; TDFA compiler.go:602-655).  n_limit < 0: all; 0: none.
 *      Writes up to cap_matches offset records; returns the number of matches the reference
 *      would return (if > cap_matches only the first cap_matches were written).
 *      The TDFA variant reproduces the reference's stride rule, so a match can repeat
 *      (SURVEY.md Q2); repeats are written out as individual records here.                      */
int64_t rgx_find_all(rgx_ctx* c, const rgx_program* p, const uint8_t* buf, uint64_t len,
                     int64_t n_limit, int64_t* out_offsets, uint64_t cap_matches);

/* Host buffers, run-length result: record j (out_offsets[j*2(k+1)..]) is one distinct match and
 * out_reps[j] how many consecutive times FindAllBytes returns it (always 1 for the backtracking
 * engine; > 1 only under the TDFA stride rule).  A cgo shim appends the same *TBytesResult
 * out_reps[j] times.  Return value / n_records / RGX_ECAPACITY as for rgx_find_all_dev.          */
int64_t rgx_find_all_rle(rgx_ctx* c, const rgx_program* p, const uint8_t* buf, uint64_t len,
                         int64_t n_limit, int64_t* out_offsets, uint32_t* out_reps,
                         uint64_t cap_records, uint64_t* n_records);

/* Device-resident form.  Results stay on the device in run-length form: record j is the offset
 * record of one distinct match and d_reps[j] how many consecutive times the reference returns
 * it (always 1 for the backtracking engine).  *n_records = distinct records written (<= cap),
 * return value = sum of reps = len(FindAllBytes(...)).  If the distinct records exceed
 * cap_records, returns RGX_ECAPACITY and sets *n_records to the number required.             */
int64_t rgx_find_all_dev(rgx_ctx* c, const rgx_program* p, const uint8_t* d_buf, uint64_t len,
                         int64_t n_limit, int64_t* d_out_offsets, uint32_t* d_reps,
                         uint64_t cap_records, uint64_t* n_records);

/* One shard of a logical buffer that is split over several GPUs.  d_buf holds shard_len bytes of
 * this rank's shard followed by (buf_len - shard_len) halo bytes (the start of the next rank's
 * shard; none on the last rank; is_last = d_buf ends where the logical input ends).  Only match starts inside the shard are reported; offsets are
 * written as out_base + shard-relative position.  entry_cursor is where the reference's FindAll
 * cursor stands when it reaches this shard (shard-relative; rank 0: 0; otherwise the previous rank's
 * exit_cursor minus this shard's out_base, or a guess that is corrected later).  reuse_scan is a
 * bit set: 1 = the records of the previous call on this context are reused (no scan), 2 = replay
 * the cursor only (no output; *n_records = 0, *exit_cursor valid), 4 = output only, from the records
 * and the replay of the previous call.  A multi-GPU caller scans and replays with 2, all-gathers the
 * exit cursors, replays again with 2|1 where its entry was wrong, and ends with 4
 * (regengo_b200/dist.py).  A match attempt
 * that runs off the halo fails the call with RGX_ECAPACITY.  Only patterns on the fast TDFA scan
 * path are supported in this version (others: RGX_EUNSUPPORTED).                                  */
int64_t rgx_find_all_shard_dev(rgx_ctx* c, const rgx_program* p, const uint8_t* d_buf, uint64_t buf_len,
                               uint64_t shard_len, int32_t is_last, int64_t entry_cursor, int64_t out_base,
                               int32_t reuse_scan, int64_t* d_out_offsets, uint32_t* d_reps,
                               uint64_t cap_records, uint64_t* n_records, int64_t* exit_cursor);

/* The same with a PRE-HALO: d_buf starts pre_len bytes (a multiple of 128 KiB; d_buf 16-byte aligned) before the
 * shard.  The cursor is replayed through those bytes from a guess and nothing of them is reported; cursors entering the
 * same records merge within a few of them, so *entry_at_shard (shard-relative) is the true entry cursor unless the
 * pre-halo holds too few matches -- the caller checks it against the predecessor's exit cursor (one 16-byte all-gather
 * per step) and, if it ever differs, redoes that shard with rgx_find_all_shard_dev and the right entry. */
int64_t rgx_find_all_shard_pre_dev(rgx_ctx* c, const rgx_program* p, const uint8_t* d_buf, uint64_t buf_len,
                                   uint64_t pre_len, uint64_t shard_len, int32_t is_last, int64_t out_base,
                                   int64_t* d_out_offsets, uint32_t* d_reps, uint64_t cap_records, uint64_t* n_records,
                                   int64_t* entry_at_shard, int64_t* exit_cursor);

/* ---- ReplaceAllBytesAppend: func (T) ReplaceAllBytesAppend(input []byte, template string, buf []byte) []byte
 *      (internal/compiler/replace.go:192-273; template syntax replace/template.go:60-163: $0 $1 $12 $name ${1}
 *      ${name} $$).  Batch extension: input i is bytes[offs[i]:offs[i+1]]; its result is
 *      out_bytes[out_offs[i]:out_offs[i+1]] (out_offs has n + 1 entries, results back to back).  *out_total = the
 *      bytes all results take; if that exceeds out_cap nothing is written, out_offs is still filled and the call
 *      returns RGX_ECAPACITY (call again with a buffer of *out_total bytes).  A malformed template -- the reference
 *      panics -- returns RGX_EINVAL with the reference's message.  Per input the reference's loop runs literally:
 *      FindBytesReuse on input[offset:], bytes.Index for the text, template expansion from that match's captures;
 *      references to groups the pattern does not have expand to nothing.  Generated only for patterns with capture
 *      groups (compiler.go:310-337), like Find*.                                                              */
int rgx_replace_batch(rgx_ctx* c, const rgx_program* p, const char* tmpl, uint64_t tmpl_len, const uint8_t* bytes,
                      const uint64_t* offs, uint64_t n, uint8_t* out_bytes, uint64_t out_cap, uint64_t* out_offs,
                      uint64_t* out_total);
int rgx_replace_batch_dev(rgx_ctx* c, const rgx_program* p, const char* tmpl, uint64_t tmpl_len, const uint8_t* d_bytes,
                          const uint64_t* d_offs, uint64_t n, uint8_t* d_out_bytes, uint64_t out_cap,
                          uint64_t* d_out_offs, uint64_t* out_total);
/* replace.Parse + the lookups of the generated expansion, without running anything: RGX_OK or RGX_EINVAL.  The dump
 * is a JSON list of resolved segments [[group or -1, "literal bytes in hex"], ...] (host only; returns its length). */
int rgx_replace_template_check(const rgx_program* p, const char* tmpl, uint64_t tmpl_len);
int64_t rgx_replace_template_dump(const rgx_program* p, const char* tmpl, uint64_t tmpl_len, char* buf, size_t cap);

/* ---- FindReader: func (T) FindReader(r io.Reader, cfg stream.Config, onMatch ...) error
 *      (streaming.go:85-255) for a reader that fills every Read (bytes.Reader semantics) over
 *      `stream[0:len]`.  buffer_size/max_leftover are stream.Config{BufferSize, MaxLeftover}
 *      before ApplyDefaults (0 => defaults).  Per match: StreamOffset, ChunkIndex and the offset
 *      record in ABSOLUTE STREAM OFFSETS (the reference's slices alias its chunk buffer; a shim
 *      rebuilds them as stream[s:e] or subtracts the chunk's stream offset).  Early stop is the caller's
 *      (truncate at the first `false`).  Returns the match count (see rgx_find_all for cap).
 *      first_chunk/n_chunks select a chunk range [first_chunk, first_chunk+n_chunks) so that a
 *      stream can be sharded over GPUs (n_chunks < 0: to the end).                              */
int64_t rgx_find_reader(rgx_ctx* c, const rgx_program* p, const uint8_t* stream, uint64_t len,
                        int64_t buffer_size, int64_t max_leftover, int64_t first_chunk,
                        int64_t n_chunks, int64_t* out_stream_off, int32_t* out_chunk_idx,
                        int64_t* out_offsets, uint64_t cap_matches);
/* Device-resident stream; d_stream points at stream byte `base` (the shard's first byte) and
 * holds `len` bytes of it; total_len is the length of the whole stream.                         */
int64_t rgx_find_reader_dev(rgx_ctx* c, const rgx_program* p, const uint8_t* d_stream,
                            uint64_t base, uint64_t len, uint64_t total_len, int64_t buffer_size,
                            int64_t max_leftover, int64_t first_chunk, int64_t n_chunks,
                            int64_t* d_out_stream_off, int32_t* d_out_chunk_idx,
                            int64_t* d_out_offsets, uint64_t cap_matches);
/* stream.Config.Validate + ApplyDefaults (stream/stream.go:96-134) for this pattern. */
int rgx_stream_config(const rgx_program* p, int64_t buffer_size, int64_t max_leftover,
                      int64_t* eff_buffer_size, int64_t* eff_max_leftover);

/* ---- device memory helpers (so that a host language without CUDA bindings can stage data) -- */
int rgx_dev_alloc(rgx_ctx* c, size_t bytes, void** out);
int rgx_dev_free(rgx_ctx* c, void* p);
int rgx_dev_upload(rgx_ctx* c, void* d_dst, const void* h_src, size_t bytes);
int rgx_dev_download(rgx_ctx* c, void* h_dst, const void* d_src, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* REGENGO_B200_H */
