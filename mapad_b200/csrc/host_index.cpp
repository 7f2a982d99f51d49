// host_index.cpp — host side of the index: the content of mapAD's index files as in-memory arrays.
// Mirrors `mapad index` (/root/reference/src/index/indexing.rs:29-256) and the loaders of
// src/index/mod.rs:212-239.  Suffix sorting is our own SA-IS (sais.hpp).
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/mapad_gpu.h"
#include "host_index.hpp"
#include "sais.hpp"

namespace mapad {

GpuSuffixSortFn g_gpu_suffix_sort = nullptr;

static inline uint8_t complement_sym(uint8_t b) {
  switch (b) {
    case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C';
    default: return b;  // 'X', '$'
  }
}

// IUPAC symbol -> candidate bases (src/index/mod.rs:16-29, indexing.rs:79-93)
static const char* iupac_choices(uint8_t s) {
  switch (s) {
    case 'R': return "AG"; case 'Y': return "CT"; case 'K': return "GT"; case 'M': return "AC";
    case 'S': return "CG"; case 'W': return "AT"; case 'B': return "CGT"; case 'D': return "AGT";
    case 'H': return "ACT"; case 'V': return "ACG"; case 'N': return "ACGT"; case 'U': return "T";
    default: return nullptr;
  }
}

struct SplitMix64 {
  uint64_t s;
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
};

void HostIndex::refresh_view() {
  memset(&view, 0, sizeof view);
  view.n = n;
  view.bwt = bwt.data();
  for (int i = 0; i < 8; ++i) view.less[i] = less[i];
  view.sentinel_rows[0] = sentinel_rows[0];
  view.sentinel_rows[1] = sentinel_rows[1];
  view.sa_sample = sa_sample.data();
  view.n_sa_samples = sa_sample.size();
  view.sa_rate = sa_rate;
  view.extra_rows = extra_rows.data();
  view.n_extra_rows = extra_rows.size() / 2;
  view.n_contigs = contig_start.size();
  view.contig_start = contig_start.data();
  view.contig_end = contig_end.data();
  name_ptrs.clear();
  for (const std::string& s : contig_names) name_ptrs.push_back(s.c_str());
  view.contig_name = name_ptrs.data();
  view.orig_pos = orig_pos.data();
  view.orig_sym = orig_sym.data();
  view.n_orig = orig_pos.size();
}

void HostIndex::derive_from_bwt() {
  for (int i = 0; i < 8; ++i) less[i] = 0;
  uint64_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int k = 0;
  sentinel_rows[0] = sentinel_rows[1] = 0;
  for (uint64_t i = 0; i < n; ++i) {
    uint8_t c = bwt[i];
    cnt[c & 7]++;
    if (c == 0 && k < 2) sentinel_rows[k++] = i;
  }
  uint64_t acc = 0;
  for (int c = 0; c < 8; ++c) { less[c] = acc; acc += cnt[c]; }
}

template <class Int>
static void build_sa_bwt(HostIndex& ix, const std::vector<uint8_t>& ranks) {
  const uint64_t n = ranks.size();
  // rust-bio transform_text: the last sentinel is the smallest symbol, an earlier one the next
  std::vector<uint8_t> t(n);
  for (uint64_t i = 0; i < n; ++i) t[i] = ranks[i] == 0 ? (i == n - 1 ? 0 : 1) : (uint8_t)(ranks[i] + 1);
  std::vector<Int> sa(n);
  Sais<Int, uint8_t>::build(t.data(), sa.data(), (Int)n, (Int)8);
  t.clear();
  t.shrink_to_fit();
  ix.bwt.resize(n);
  ix.sa_sample.clear();
  ix.extra_rows.clear();
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t p = (uint64_t)sa[i];
    uint8_t c = p > 0 ? ranks[p - 1] : ranks[n - 1];
    ix.bwt[i] = c;
    if (i % ix.sa_rate == 0) ix.sa_sample.push_back(p);
    else if (c == 0) { ix.extra_rows.push_back(i); ix.extra_rows.push_back(p); }  // index/mod.rs:113-118
  }
}

int HostIndex::build(uint64_t n_contigs, const char* const* names, const char* const* seqs, const uint64_t* lens,
                     uint64_t seed, const char* draws, uint64_t n_draws, int gpu_device) {
  std::vector<uint8_t> ref;
  uint64_t total = 0;
  for (uint64_t c = 0; c < n_contigs; ++c) total += lens[c];
  ref.reserve(total);
  contig_start.clear(); contig_end.clear(); contig_names.clear();
  uint64_t end = 0;
  for (uint64_t c = 0; c < n_contigs; ++c) {
    for (uint64_t i = 0; i < lens[c]; ++i) {
      uint8_t b = (uint8_t)std::toupper((unsigned char)seqs[c][i]);
      bool ok = b == 'A' || b == 'C' || b == 'G' || b == 'T' || iupac_choices(b) != nullptr;
      if (!ok) return MAPAD_EINVAL;  // Error::ParseError (indexing.rs:69-75)
      ref.push_back(b);
    }
    end += lens[c];
    contig_start.push_back(end - lens[c]);
    contig_end.push_back(end - 1);
    contig_names.push_back(names && names[c] ? names[c] : "");
  }
  // run_apply (indexing.rs:217-256)
  orig_pos.clear(); orig_sym.clear();
  SplitMix64 rng{seed};
  uint64_t draw_i = 0;
  for (uint64_t i = 0; i < ref.size();) {
    uint8_t sym = ref[i];
    uint64_t run = 1;
    while (i + run < ref.size() && ref[i + run] == sym) ++run;
    if (!(sym == 'A' || sym == 'C' || sym == 'G' || sym == 'T')) {
      if (run < 20) {
        const char* ch = iupac_choices(sym);
        size_t nch = strlen(ch);
        for (uint64_t j = 0; j < run; ++j) {
          uint8_t rep;
          if (nch == 1) rep = (uint8_t)ch[0];
          else if (draws && draw_i < n_draws) rep = (uint8_t)draws[draw_i++];
          else rep = (uint8_t)ch[rng.next() % nch];
          orig_pos.push_back(i + j);
          orig_sym.push_back(sym);
          ref[i + j] = rep;
        }
      } else {
        for (uint64_t j = 0; j < run; ++j) ref[i + j] = 'X';
      }
    }
    i += run;
  }
  // text = fwd $ revcomp $ (indexing.rs:139-144), rank-transformed over $ACGTX (indexing.rs:147-152)
  auto rank = [](uint8_t b) -> uint8_t {
    switch (b) { case 'A': return 1; case 'C': return 2; case 'G': return 3; case 'T': return 4; case 'X': return 5; default: return 0; }
  };
  const uint64_t G = ref.size();
  n = 2 * G + 2;
  std::vector<uint8_t> ranks(n);
  for (uint64_t i = 0; i < G; ++i) ranks[i] = rank(ref[i]);
  ranks[G] = 0;
  for (uint64_t i = 0; i < G; ++i) ranks[G + 1 + i] = rank(complement_sym(ref[G - 1 - i]));
  ranks[n - 1] = 0;
  ref.clear();
  ref.shrink_to_fit();
  sa_rate = 32;
  bool sorted_on_device = false;
  if (gpu_device >= 0) {
    if (!g_gpu_suffix_sort) return MAPAD_ENODEV;
    int rc = g_gpu_suffix_sort(ranks, gpu_device, *this);
    if (rc == MAPAD_OK) sorted_on_device = true;
    else if (rc != MAPAD_EINDEX) return rc;
  }
  if (!sorted_on_device) {
    if (n < (1ull << 31)) build_sa_bwt<int32_t>(*this, ranks);
    else build_sa_bwt<int64_t>(*this, ranks);
  }
  derive_from_bwt();
  refresh_view();
  return MAPAD_OK;
}

int HostIndex::from_view(const mapad_index_view& v) {
  n = v.n;
  bwt.assign(v.bwt, v.bwt + v.n);
  sa_sample.assign(v.sa_sample, v.sa_sample + v.n_sa_samples);
  sa_rate = v.sa_rate;
  extra_rows.assign(v.extra_rows, v.extra_rows + 2 * v.n_extra_rows);
  contig_start.assign(v.contig_start, v.contig_start + v.n_contigs);
  contig_end.assign(v.contig_end, v.contig_end + v.n_contigs);
  contig_names.clear();
  for (uint64_t i = 0; i < v.n_contigs; ++i) contig_names.push_back(v.contig_name && v.contig_name[i] ? v.contig_name[i] : "");
  orig_pos.assign(v.orig_pos, v.orig_pos + v.n_orig);
  orig_sym.assign(v.orig_sym, v.orig_sym + v.n_orig);
  derive_from_bwt();
  if (sa_rate == 0 || sa_sample.size() != (n + sa_rate - 1) / sa_rate) return MAPAD_EINDEX;
  refresh_view();
  return MAPAD_OK;
}

}  // namespace mapad

using mapad::HostIndex;

extern "C" {

static int build_common(uint64_t n_contigs, const char* const* names, const char* const* sequences,
                        const uint64_t* lengths, uint64_t seed, const char* draws, uint64_t n_draws, mapad_index** out,
                        int gpu_device = -1) {
  if (!out || !sequences || !lengths) return MAPAD_EINVAL;
  *out = nullptr;
  HostIndex* ix = new (std::nothrow) HostIndex();
  if (!ix) return MAPAD_ENOMEM;
  int rc;
  try { rc = ix->build(n_contigs, names, sequences, lengths, seed, draws, n_draws, gpu_device); }
  catch (const std::bad_alloc&) { rc = MAPAD_ENOMEM; }
  catch (...) { rc = MAPAD_EINVAL; }
  if (rc != MAPAD_OK) { delete ix; return rc; }
  *out = reinterpret_cast<mapad_index*>(ix);
  return MAPAD_OK;
}

int mapad_index_build(uint64_t n_contigs, const char* const* names, const char* const* sequences, const uint64_t* lengths,
                      uint64_t seed, mapad_index** out) {
  return build_common(n_contigs, names, sequences, lengths, seed, nullptr, 0, out);
}

int mapad_index_build_on_device(uint64_t n_contigs, const char* const* names, const char* const* sequences,
                                const uint64_t* lengths, uint64_t seed, int device, mapad_index** out) {
  if (device < 0) return MAPAD_EINVAL;
  return build_common(n_contigs, names, sequences, lengths, seed, nullptr, 0, out, device);
}

int mapad_index_build_with_draws(uint64_t n_contigs, const char* const* names, const char* const* sequences,
                                 const uint64_t* lengths, const char* replacement_draws, uint64_t n_draws,
                                 mapad_index** out) {
  return build_common(n_contigs, names, sequences, lengths, 0, replacement_draws, n_draws, out);
}

int mapad_index_from_view(const mapad_index_view* v, mapad_index** out) {
  if (!v || !out || !v->bwt || !v->sa_sample) return MAPAD_EINVAL;
  *out = nullptr;
  HostIndex* ix = new (std::nothrow) HostIndex();
  if (!ix) return MAPAD_ENOMEM;
  int rc;
  try { rc = ix->from_view(*v); }
  catch (const std::bad_alloc&) { rc = MAPAD_ENOMEM; }
  catch (...) { rc = MAPAD_EINVAL; }
  if (rc != MAPAD_OK) { delete ix; return rc; }
  *out = reinterpret_cast<mapad_index*>(ix);
  return MAPAD_OK;
}

int mapad_index_get_view(const mapad_index* ix, mapad_index_view* out) {
  if (!ix || !out) return MAPAD_EINVAL;
  *out = reinterpret_cast<const HostIndex*>(ix)->view;
  return MAPAD_OK;
}

void mapad_index_free(mapad_index* ix) { delete reinterpret_cast<HostIndex*>(ix); }

}  // extern "C"
