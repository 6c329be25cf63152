// Internal definitions shared by the two halves of the C ABI.
#pragma once
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/regengo_b200.h"
#include "blob.hpp"
#include "frontend/program.hpp"

namespace rgx {
void set_error(const std::string& s);
struct DeviceImage;  // packed tables resident on one device (device_program.cuh)
}

struct rgx_program {
  rgx::Program prog;
  std::vector<uint32_t> blob;
  std::string json;
  // lazily built per-device images of the packed program (immutable once built)
  mutable std::mutex mu;
  mutable std::vector<rgx::DeviceImage*> images;  // indexed by device ordinal
};

namespace rgx {
// implemented in capi_device.cu; frees the per-device images of p
void release_device_program(rgx_program* p);
}
