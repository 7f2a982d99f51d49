// tests/emu/simt_emu.hpp — TEST INFRASTRUCTURE.  A minimal SIMT emulator: the lanes of the lane groups of
// mapad_b200/csrc/search_group.cuh run as cooperatively scheduled coroutines on one OS thread; the group collectives
// (mapad_simt_emu_shfl / _ballot / _sync, declared in csrc/simt.cuh) are rendezvous points of the lanes of one group.
// This lets the non-GPU test-suite execute the very source of the cooperative kernel, with divergent control flow
// between groups, and compare it bit for bit with the oracle.
#pragma once
#include <cstdint>
#include <functional>
#include <vector>

namespace simt_emu {

// Runs fn(group, lane_in_group) for every lane of `n_groups` groups of `group_size` lanes until all have returned.
// Lanes are scheduled round robin; a lane runs until it reaches a collective whose partners have not arrived yet.
void run(int n_groups, int group_size, const std::function<void(int, int)>& fn);

}  // namespace simt_emu
