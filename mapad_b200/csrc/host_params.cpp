// host_params.cpp — host-side mirror of the reference's scoring traits: SequenceDifferenceModel
// (/root/reference/src/map/sequence_difference_models.rs:14-62,104-419) and MismatchBound
// (src/map/mismatch_bounds.rs:10-281), plus the derivation of AlignmentParameters from CLI flags
// (src/main.rs:418-499).  Uses the host's libm exactly like the reference binary does; the device
// re-evaluates the built-in aDNA model with the bit-exact restatements in libm_emu.cuh.
#include <cmath>
#include <cstring>
#include <limits>

#include "../../include/mapad_gpu.h"
#include "host_params.hpp"

namespace mapad {

static inline float powi_rt(float a, int b) {  // compiler-rt __powisf2 == Rust f32::powi
  const bool recip = b < 0;
  float r = 1.0f;
  while (true) {
    if (b & 1) r *= a;
    b /= 2;
    if (b == 0) break;
    a *= a;
  }
  return recip ? 1.0f / r : r;
}

float qual2prob(uint8_t q) { return powf(10.0f, -(float)q / 10.0f) / 3.0f; }  // :275-277

static float simple_get(const mapad_params& p, size_t i, size_t read_length, uint8_t from, uint8_t to, uint8_t q) {  // :117-207
  const size_t fp_dist = i, tp_dist = read_length - 1 - i;
  const float seq_err = p.ignore_base_quality ? qual2prob(255) : qual2prob(q);
  const float indep = fmaf(seq_err, -p.divergence, seq_err + p.divergence);
  float c_to_t = 0.f, g_to_a = 0.f;
  if ((from == 'C' && (to == 'C' || to == 'T')) || (from == 'G' && (to == 'A' || to == 'G'))) {
    float p_fwd, p_rev;
    if (p.library == MAPAD_LIB_SINGLE_STRANDED) {
      float a = powi_rt(p.five_prime_overhang, (int)fp_dist + 1);
      float b = powi_rt(p.three_prime_overhang, (int)tp_dist + 1);
      p_fwd = fmaf(a, -b, a + b);
      p_rev = 0.0f;
    } else {
      p_fwd = powi_rt(p.five_prime_overhang, (int)fp_dist + 1);
      p_rev = powi_rt(p.five_prime_overhang, (int)tp_dist + 1);
    }
    c_to_t = fmaf(p.ss_deamination_rate, p_fwd, p.ds_deamination_rate * (1.0f - p_fwd));
    g_to_a = fmaf(p.ss_deamination_rate, p_rev, p.ds_deamination_rate * (1.0f - p_rev));
  }
  float v;
  switch (from) {
    case 'A': v = to == 'A' ? fmaf(3.0f, -indep, 1.0f) : indep; break;
    case 'C':
      if (to == 'C') v = fmaf(4.0f * indep, c_to_t, fmaf(3.0f, -indep, 1.0f) - c_to_t);
      else if (to == 'T') v = fmaf(4.0f * indep, -c_to_t, indep + c_to_t);
      else v = indep;
      break;
    case 'G':
      if (to == 'A') v = fmaf(4.0f * indep, -g_to_a, indep + g_to_a);
      else if (to == 'G') v = fmaf(4.0f * indep, g_to_a, fmaf(3.0f, -indep, 1.0f) - g_to_a);
      else v = indep;
      break;
    case 'T': v = to == 'T' ? fmaf(3.0f, -indep, 1.0f) : indep; break;
    default: v = indep;
  }
  return log2f(fmaxf(v, std::numeric_limits<float>::epsilon()));
}

static float vindija_get(size_t i, size_t read_length, uint8_t from, uint8_t to) {  // :353-394
  static const float ppm[7] = {0.4f, 0.25f, 0.1f, 0.06f, 0.05f, 0.04f, 0.03f};
  float p;
  if (from == 'C') {
    size_t k = i < read_length - (i + 1) ? i : read_length - (i + 1);
    float pct = k < 7 ? ppm[k] : 0.02f;
    p = to == 'T' ? pct : (to == 'C' ? 1.0f - pct : 0.0005f);
  } else {
    p = from == to ? 1.0f - 0.0005f : 0.0005f;
  }
  return log2f(p);
}

float sdm_get(const mapad_params& p, size_t i, size_t read_length, uint8_t from, uint8_t to, uint8_t q) {
  switch (p.model_kind) {
    case MAPAD_MODEL_SIMPLE_ADNA: return simple_get(p, i, read_length, from, to, q);
    case MAPAD_MODEL_VINDIJA_PWM: return vindija_get(i, read_length, from, to);
    case MAPAD_MODEL_TEST:  // :409-419
      if (from == 'C' && to == 'T') return p.test_deam_score;
      return from == to ? p.test_match_score : p.test_mm_score;
    default:
      return p.custom_get ? p.custom_get(p.custom_user, i, read_length, from, to, q) : 0.0f;
  }
}

float sdm_repr_mm(const mapad_params& p) {  // :16-31
  return sdm_get(p, 40, 80, 'T', 'A', 255) - sdm_get(p, 40, 80, 'T', 'T', 255);
}

int16_t sdm_alignment_start(const mapad_params& p, size_t len) {  // :59-61, :209-211
  if (p.model_kind == MAPAD_MODEL_SIMPLE_ADNA) return (int16_t)len;
  if (p.model_kind == MAPAD_MODEL_CUSTOM && p.custom_start) return p.custom_start(p.custom_user, len);
  return (int16_t)((int16_t)len / 2);
}

float discrete_max_mismatches(size_t read_length, float thr, float rate) {  // mismatch_bounds.rs:209-236
  float lambda = (float)read_length * rate;
  float eml = expf(-lambda);
  float sum = eml;
  if (!(1.0f - sum > thr)) return 0.0f;
  uint64_t last_k = 1;
  float lk = 1.0f;
  uint64_t kf = 1;
  for (uint64_t k = 1; k <= (uint64_t)read_length; ++k) {
    lk *= lambda;
    kf *= k;
    sum += lk * eml / (float)kf;
    if (1.0f - sum > thr) last_k = k + 1; else break;
  }
  return (float)last_k;
}

float bound_table_value(const mapad_params& p, size_t read_length) {
  if (p.bound_kind == MAPAD_BOUND_DISCRETE) {  // Discrete::get (:238-255)
    if (read_length < 17) return 0.0f;
    return discrete_max_mismatches(read_length, p.poisson_threshold, p.base_error_rate);
  }
  if (p.bound_kind == MAPAD_BOUND_CONTINUOUS) return powf((float)read_length, p.exponent);  // :116-121
  return 0.0f;
}

void finalize_params(const mapad_params& in, DevParams& d, float qual_table[256]) {
  memset(&d, 0, sizeof d);
  d.model = in.model_kind == MAPAD_MODEL_SIMPLE_ADNA ? MODEL_SIMPLE : MODEL_TABLE;
  d.library = in.library;
  d.overhang5 = in.five_prime_overhang;
  d.overhang3 = in.library == MAPAD_LIB_DOUBLE_STRANDED ? in.five_prime_overhang : in.three_prime_overhang;
  d.ds_rate = in.ds_deamination_rate;
  d.ss_rate = in.ss_deamination_rate;
  d.divergence = in.divergence;
  d.ignore_q = in.ignore_base_quality;
  d.default_q_prob = qual2prob(255);
  for (int q = 0; q < 256; ++q) qual_table[q] = qual2prob((uint8_t)q);
  d.start_mode = in.model_kind == MAPAD_MODEL_SIMPLE_ADNA ? 0 : (in.model_kind == MAPAD_MODEL_CUSTOM && in.custom_start ? 2 : 1);
  d.bound_kind = in.bound_kind;
  d.repr_mm = in.representative_mismatch_penalty != 0.0f ? in.representative_mismatch_penalty : sdm_repr_mm(in);
  d.cutoff = in.cutoff;
  d.test_threshold = in.test_threshold;
  d.test_repr_mm = in.test_representative_mm;
  d.gap_open = in.penalty_gap_open;
  d.gap_extend = in.penalty_gap_extend;
  d.gap_dist_ends = in.gap_dist_ends;
  d.max_num_gaps_open = in.max_num_gaps_open;
  d.stack_limit_abort = in.stack_limit_abort;
  d.stack_limit = in.stack_limit ? in.stack_limit : 2000000u;
  d.edit_tree_limit = in.edit_tree_limit ? in.edit_tree_limit : 10000000u;
}

int prepare_batch(const mapad_params& p, const mapad_reads& in, BatchPrep& out) {
  finalize_params(p, out.dp, out.qual_table);
  if (in.n_reads && (!in.seq || !in.qual || !in.offsets)) return MAPAD_EINVAL;
  out.max_len = 0;
  out.total_bases = in.n_reads ? in.offsets[in.n_reads] - in.offsets[0] : 0;
  for (uint64_t r = 0; r < in.n_reads; ++r) {
    if (in.offsets[r + 1] < in.offsets[r]) return MAPAD_EINVAL;
    uint64_t L = in.offsets[r + 1] - in.offsets[r];
    if (L > 32767) return MAPAD_EINVAL;  // Error::SeqLenError: reads can not be longer than i16::MAX (record.rs:145)
    if (L > out.max_len) out.max_len = (uint32_t)L;
  }
  out.bound_table.assign((size_t)out.max_len + 1, 0.0f);
  for (uint32_t L = 0; L <= out.max_len; ++L) out.bound_table[L] = bound_table_value(p, L);
  out.dp.bound_table_len = out.max_len + 1;
  out.starts.clear();
  if (out.dp.start_mode == 2) {
    out.starts.resize(in.n_reads);
    for (uint64_t r = 0; r < in.n_reads; ++r) out.starts[r] = sdm_alignment_start(p, in.offsets[r + 1] - in.offsets[r]);
  }
  out.custom_pen.clear();
  if (out.dp.model == MODEL_TABLE && !in.custom_penalties) {
    static const uint8_t ACGT[4] = {'A', 'C', 'G', 'T'};
    const uint64_t base0 = in.n_reads ? in.offsets[0] : 0;
    out.custom_pen.resize(4 * out.total_bases);
    for (uint64_t r = 0; r < in.n_reads; ++r) {
      const uint64_t o = in.offsets[r], L = in.offsets[r + 1] - o;
      for (uint64_t j = 0; j < L; ++j)
        for (int b = 0; b < 4; ++b)
          out.custom_pen[4 * (o - base0 + j) + b] = sdm_get(p, j, L, ACGT[b], in.seq[o + j], in.qual[o + j]);
    }
  }
  return MAPAD_OK;
}

}  // namespace mapad

extern "C" {

float mapad_sdm_get(const mapad_params* p, size_t i, size_t read_length, uint8_t from, uint8_t to, uint8_t q) {
  return p ? mapad::sdm_get(*p, i, read_length, from, to, q) : 0.0f;
}
float mapad_sdm_representative_mismatch_penalty(const mapad_params* p) { return p ? mapad::sdm_repr_mm(*p) : 0.0f; }
float mapad_bound_allowed_mismatches(const mapad_params* p, size_t read_length) {
  if (!p || p->bound_kind != MAPAD_BOUND_DISCRETE) return 0.0f;
  return mapad::bound_table_value(*p, read_length);
}

int mapad_params_from_cli(mapad_params* p, const char* library, float poisson_prob, float f, float t, float d, float s,
                          float divergence, float indel_rate, float gap_extension_fraction, uint8_t gap_dist_ends,
                          uint8_t max_num_gaps_open, int ignore_base_quality, int no_search_limit_recovery) {
  if (!p || !library) return MAPAD_EINVAL;
  memset(p, 0, sizeof *p);
  p->model_kind = MAPAD_MODEL_SIMPLE_ADNA;
  if (!strcmp(library, "single_stranded")) p->library = MAPAD_LIB_SINGLE_STRANDED;
  else if (!strcmp(library, "double_stranded")) p->library = MAPAD_LIB_DOUBLE_STRANDED;
  else return MAPAD_EINVAL;
  p->five_prime_overhang = f;
  p->three_prime_overhang = t;
  p->ds_deamination_rate = d;
  p->ss_deamination_rate = s;
  p->divergence = divergence / 3.0f;  // main.rs:452
  p->ignore_base_quality = ignore_base_quality;
  p->bound_kind = MAPAD_BOUND_DISCRETE;  // main.rs:456-462
  p->poisson_threshold = poisson_prob;
  p->base_error_rate = divergence;
  const float repr = mapad::sdm_repr_mm(*p);
  p->representative_mismatch_penalty = repr;
  p->penalty_gap_open = log2f(indel_rate);               // main.rs:478-481
  p->penalty_gap_extend = gap_extension_fraction * repr;  // main.rs:482-485
  p->gap_dist_ends = gap_dist_ends;
  p->max_num_gaps_open = max_num_gaps_open;
  p->stack_limit_abort = no_search_limit_recovery ? 1 : 0;
  return MAPAD_OK;
}

int mapad_abi_version(void) { return MAPAD_ABI_VERSION; }
uint64_t mapad_abi_sizeof(int what) {
  switch (what) {
    case 0: return sizeof(mapad_params); case 1: return sizeof(mapad_reads); case 2: return sizeof(mapad_edit_op);
    case 3: return sizeof(mapad_hit); case 4: return sizeof(mapad_alt); case 5: return sizeof(mapad_record);
    case 6: return sizeof(mapad_results); case 7: return sizeof(mapad_index_view); default: return 0;
  }
}

}  // extern "C"
