// host_params.hpp — see host_params.cpp
#pragma once
#include <cstddef>
#include <cstdint>

#include "../../include/mapad_gpu.h"
#include "common.h"

namespace mapad {
float qual2prob(uint8_t q);
float sdm_get(const mapad_params& p, size_t i, size_t read_length, uint8_t from, uint8_t to, uint8_t q);
float sdm_repr_mm(const mapad_params& p);
int16_t sdm_alignment_start(const mapad_params& p, size_t len);
float discrete_max_mismatches(size_t read_length, float thr, float rate);
float bound_table_value(const mapad_params& p, size_t read_length);
void finalize_params(const mapad_params& in, DevParams& d, float qual_table[256]);
}  // namespace mapad

#include <vector>
namespace mapad {
// Host-side per-batch preparation shared by every launcher of the kernels.
struct BatchPrep {
  DevParams dp;
  float qual_table[256];
  std::vector<float> bound_table;   // indexed by read length, 0..max_len
  std::vector<int16_t> starts;      // start_mode 2 only
  std::vector<float> custom_pen;    // MODEL_TABLE only, when the caller did not supply penalties
  uint32_t max_len = 0;
  uint64_t total_bases = 0;
};
int prepare_batch(const mapad_params& p, const mapad_reads& in, BatchPrep& out);
}  // namespace mapad
