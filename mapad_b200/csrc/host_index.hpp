// host_index.hpp — in-memory form of mapAD's index files (see host_index.cpp).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/mapad_gpu.h"

namespace mapad {

struct HostIndex {
  uint64_t n = 0;
  std::vector<uint8_t> bwt;            // .tbw
  uint64_t less[8] = {0};              // .tle
  uint64_t sentinel_rows[2] = {0, 0};  // RtFmdIndex::sentinel_occ
  std::vector<uint64_t> sa_sample;     // .tsa
  uint64_t sa_rate = 32;
  std::vector<uint64_t> extra_rows;    // (row, pos) pairs sorted by row
  std::vector<uint64_t> contig_start, contig_end;  // .tpi
  std::vector<std::string> contig_names;
  std::vector<uint64_t> orig_pos;      // .tos
  std::vector<uint8_t> orig_sym;
  std::vector<const char*> name_ptrs;
  mapad_index_view view;

  // gpu_device >= 0: suffix sorting on that CUDA device (gpu_index_build.cu), falling back to the host SA-IS
  // when the text is too repetitive for the device sorter
  int build(uint64_t n_contigs, const char* const* names, const char* const* seqs, const uint64_t* lens, uint64_t seed,
            const char* draws, uint64_t n_draws, int gpu_device = -1);
  int from_view(const mapad_index_view& v);
  void derive_from_bwt();
  void refresh_view();
};

typedef int (*GpuSuffixSortFn)(const std::vector<uint8_t>& ranks, int device, HostIndex& ix);
extern GpuSuffixSortFn g_gpu_suffix_sort;  // installed by gpu_index_build.cu when it is linked in

}  // namespace mapad
