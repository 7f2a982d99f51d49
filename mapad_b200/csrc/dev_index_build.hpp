// dev_index_build.hpp — re-layout of the host index into the device blob (see common.h).
#pragma once
#include <cstdint>
#include <vector>

#include "common.h"
#include "host_index.hpp"

namespace mapad {
// Fills `meta` and the blob bytes; `blob` can be uploaded verbatim (all offsets are 64 B aligned).
// layout: -1 = by size (wide iff n >= 2^31), 0 = narrow, 1 = wide (tests force the wide layout on small indexes).
int build_device_blob(const HostIndex& ix, IndexMeta& meta, std::vector<uint8_t>& blob, int layout = -1);
}  // namespace mapad
