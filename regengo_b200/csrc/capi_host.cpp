// Host-only half of the C ABI (include/regengo_b200.h): pattern compilation, inspection, blob
// round trip, stream.Config arithmetic.  Works without a GPU.  The compute entry points live in
// capi_device.cu.
#include <cstring>
#include <string>

#include "capi_internal.hpp"

namespace rgx {
thread_local std::string g_last_error;
void set_error(const std::string& s) { g_last_error = s; }
}  // namespace rgx

using namespace rgx;

extern "C" {

const char* rgx_last_error(void) { return g_last_error.c_str(); }
const char* rgx_version(void) { return "regengo_b200 0.1 (sm_100a)"; }

static void finish_program(rgx_program* p) {
  p->blob = program_to_blob(p->prog);
  p->json = program_to_json(p->prog);
}

int rgx_compile(const char* pattern, const rgx_options* opts, rgx_program** out) {
  if (!pattern || !out) { set_error("rgx_compile: null argument"); return RGX_EINVAL; }
  Options o;
  if (opts) {
    o.force_thompson = opts->force_thompson != 0; o.force_tnfa = opts->force_tnfa != 0;
    o.force_tdfa = opts->force_tdfa != 0; o.tdfa_threshold = opts->tdfa_threshold;
  }
  auto* p = new rgx_program();
  std::string err;
  if (!build_program(pattern, o, p->prog, err)) {
    delete p;
    set_error("rgx_compile: " + err);
    return RGX_EPATTERN;
  }
  finish_program(p);
  *out = p;
  return RGX_OK;
}

int rgx_load(const void* blob, size_t n, rgx_program** out) {
  if (!blob || !out || n % 4 != 0) { set_error("rgx_load: bad argument"); return RGX_EINVAL; }
  auto* p = new rgx_program();
  std::string err;
  std::vector<uint32_t> w(n / 4);
  std::memcpy(w.data(), blob, n);
  if (!blob_to_program(w.data(), w.size(), p->prog, err)) {
    delete p;
    set_error("rgx_load: " + err);
    return RGX_EINVAL;
  }
  p->blob = std::move(w);
  p->json = program_to_json(p->prog);
  *out = p;
  return RGX_OK;
}

void rgx_program_free(rgx_program* p) {
  if (!p) return;
  release_device_program(p);
  delete p;
}

int rgx_program_info(const rgx_program* p, rgx_info* o) {
  if (!p || !o) { set_error("rgx_program_info: null argument"); return RGX_EINVAL; }
  std::memset(o, 0, sizeof(*o));
  const Program& P = p->prog;
  o->n_inst = (int32_t)P.prog.inst.size();
  o->num_cap = P.prog.num_cap;
  o->n_groups = P.prog.num_cap / 2 - 1;
  o->match_engine = P.match_engine;
  o->find_engine = P.find_engine;
  o->match_memo = P.match_memo;
  o->find_memo = P.find_memo;
  o->per_capture_ckpt = P.per_capture_ckpt;
  o->anchored = P.anchored;
  o->min_match_len = P.min_match_len;
  o->max_match_len = P.max_match_len;
  o->default_max_leftover = P.default_max_leftover;
  o->min_buffer = P.min_buffer;
  o->tdfa_states = P.tdfa.built && P.find_engine == FIND_TDFA ? P.tdfa.num_states : 0;
  o->tdfa_tags = P.tdfa.built && P.find_engine == FIND_TDFA ? P.tdfa.num_tags : 0;
  return RGX_OK;
}

const char* rgx_program_json(const rgx_program* p) { return p ? p->json.c_str() : nullptr; }

const char* rgx_program_group_name(const rgx_program* p, int32_t i) {
  if (!p || i < 1 || (size_t)i >= p->prog.capture_names.size()) return nullptr;
  return p->prog.capture_names[i].c_str();
}

int64_t rgx_program_blob(const rgx_program* p, void* buf, size_t cap) {
  if (!p) { set_error("rgx_program_blob: null program"); return RGX_EINVAL; }
  size_t bytes = p->blob.size() * 4;
  if (buf && cap >= bytes) std::memcpy(buf, p->blob.data(), bytes);
  return (int64_t)bytes;
}

// stream.Config.Validate + ApplyDefaults (stream/stream.go:96-134) with the pattern's constants
// (streaming.go:41-46, 56-62).  MaxLeftover == -1 panics in the reference: rejected.
int rgx_stream_config(const rgx_program* p, int64_t buffer_size, int64_t max_leftover, int64_t* eb, int64_t* el) {
  if (!p || !eb || !el) { set_error("rgx_stream_config: null argument"); return RGX_EINVAL; }
  if (buffer_size < 0 || max_leftover < 0) { set_error("rgx_stream_config: negative BufferSize/MaxLeftover is not supported"); return RGX_EINVAL; }
  const int64_t min_buffer = p->prog.min_buffer, def_left = p->prog.default_max_leftover;
  if (buffer_size > 0 && buffer_size < min_buffer) { set_error("stream: buffer size too small"); return RGX_EBUFFER_TOO_SMALL; }
  if (buffer_size == 0) buffer_size = 64 * 1024;
  if (buffer_size < min_buffer) buffer_size = min_buffer;
  if (max_leftover == 0) max_leftover = def_left;
  if (max_leftover > buffer_size / 2) max_leftover = buffer_size / 2;
  *eb = buffer_size; *el = max_leftover;
  return RGX_OK;
}

}  // extern "C"
