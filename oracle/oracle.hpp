// oracle/oracle.hpp
//
// TEST INFRASTRUCTURE ONLY.  CPU restatement of the read-mapping hot path of mapAD 0.45.0
// (reference tree: /root/reference, all Rust).  It exists to check the CUDA path and to serve
// as the CPU baseline leg of bench.py.  Nothing under mapad_b200/ may include, link or call it.
//
// Parity status: PINNED by the reference's own known-answer tests (tests/test_oracle_golden.py
// replays src/map/mapping.rs:1401-2956, src/map/bi_d_array.rs:243-309,
// src/map/mismatch_bounds.rs:288-377, src/map/sequence_difference_models.rs:426-1339,
// src/map/prrange.rs:186-261 and tests/integration_tests.rs:464-868).  Unpinned corners are
// listed in DESIGN.md (pop_min eviction path, rand-crate driven choices, on-disk rust-bio types).
//
// Third-party semantics restated here (not vendored under /root/reference):
//   bio 1.5.0-mapAD (rust-bio fork, Cargo.lock:122-124): suffix order with two sentinels,
//     bwt, less, Occ (inclusive rank), RankTransform, dna::complement.
//   min-max-heap 1.3.1-alpha.0 (Cargo.lock:1106-1108): push / pop_max / pop_min.
//   slab 0.4.12: LIFO key reuse.
//   Rust std BinaryHeap: push / peek / pop / into_sorted_vec.
//   glibc libm float functions (log2f, powf, expf, exp2f, log10f), compiler-rt __powisf2, fmaf.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <vector>

namespace ora {

typedef uint64_t usize;

// ------------------------------------------------------------------------------------------
// Alphabet helpers (rust-bio alphabets::dna)
// ------------------------------------------------------------------------------------------
inline uint8_t complement(uint8_t b) {
  // bio::alphabets::dna::complement: IUPAC pairs, identity for everything else.
  switch (b) {
    case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C';
    case 'a': return 't'; case 't': return 'a'; case 'c': return 'g'; case 'g': return 'c';
    case 'Y': return 'R'; case 'R': return 'Y'; case 'K': return 'M'; case 'M': return 'K';
    case 'D': return 'H'; case 'H': return 'D'; case 'V': return 'B'; case 'B': return 'V';
    case 'y': return 'r'; case 'r': return 'y'; case 'k': return 'm'; case 'm': return 'k';
    case 'd': return 'h'; case 'h': return 'd'; case 'v': return 'b'; case 'b': return 'v';
    default: return b;  // W, S, N, X, $ ...
  }
}

// ------------------------------------------------------------------------------------------
// Numerics
// ------------------------------------------------------------------------------------------
// compiler-rt __powisf2, which is what Rust's f32::powi lowers to.
inline float powi_f32(float a, int b) {
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
// Rust f32::max / f32::min (IEEE maxNum/minNum; no NaNs on this path)
inline float fmax_rs(float a, float b) { return std::fmax(a, b); }
inline float fmin_rs(float a, float b) { return std::fmin(a, b); }

struct Counters {
  uint64_t frames_popped = 0;   // P   (mapping.rs:1058)
  uint64_t d_ext_steps = 0;     // E   (bi_d_array.rs:131-136, one per single-symbol extension)
  uint64_t lf_steps = 0;        // W   (index/mod.rs:165-185, LF steps + 1 sample read per position)
  uint64_t located = 0;         // positions resolved through the sampled SA
  uint64_t tree_nodes = 0;
  uint64_t max_stack = 0;
  uint64_t limit_hit = 0;
};

// ------------------------------------------------------------------------------------------
// FMD index on rank-transformed text (src/map/fmd_index.rs + rust-bio bwt/less/occ)
// ------------------------------------------------------------------------------------------
struct BiInterval {
  usize lower = 0, lower_rev = 0, size = 0;
  BiInterval swapped() const { return BiInterval{lower_rev, lower, size}; }  // fmd_index.rs:207
};

struct Contig {
  uint64_t start, end;  // inclusive end (index/mod.rs:32-36)
  std::string name;
};

struct Index {
  usize n = 0;                     // text length incl. both sentinels
  std::vector<uint8_t> bwt;        // rank bytes
  std::vector<usize> less;         // less[c] = #symbols < c ; size max_rank + 2
  uint32_t occ_k = 128;            // Occ sampling (indexing.rs:188 uses 128, utils.rs:30 uses 3)
  std::vector<std::vector<usize>> occ_cp;  // occ_cp[c][i] = #c in bwt[0..=i*k]
  usize sentinel_occ[2] = {0, 0};  // fmd_index.rs:38-47
  int rank_of[256];                // RankTransform.ranks; -1 = not in alphabet
  std::vector<uint8_t> back_transform;     // fmd_index.rs:49-54
  // sampled suffix array (index/mod.rs:80-187)
  std::vector<usize> sa_sample;
  usize sa_rate = 32;
  std::map<usize, usize> extra_rows;
  uint8_t sentinel = 0;
  std::vector<usize> full_sa;      // optional raw SA (tests use RawSuffixArray, utils.rs:26)
  // contigs + replaced symbols (index/mod.rs:32-75,198-210)
  std::vector<Contig> contigs;
  std::map<usize, uint8_t> original_symbols;

  Index() { for (int& r : rank_of) r = -1; }

  // rust-bio Occ::get: #a in bwt[0..=r] (inclusive).
  usize occ(usize r, uint8_t a) const {
    usize i = r / occ_k;
    usize count = occ_cp[a][i];
    const uint8_t* p = bwt.data();
    for (usize j = i * occ_k + 1; j <= r; ++j) count += (p[j] == a);
    return count;
  }

  BiInterval init_interval() const { return BiInterval{0, 0, n}; }  // fmd_index.rs:67-73

  // fmd_index.rs:140-146
  usize sentinel_count_upto(usize pos) const {
    for (usize i = 0; i < 2; ++i)
      if (pos < sentinel_occ[i]) return i;
    return 2;
  }

  // FmdExtIterator (fmd_index.rs:117-182): yields c = 4,3,2,1 (T,G,C,A) in that order.
  void extend_all(const BiInterval& in, BiInterval out[4]) const {
    usize o0 = in.lower == 0 ? 0 : sentinel_count_upto(in.lower - 1);
    usize s = sentinel_count_upto(in.lower + in.size - 1) - o0;
    usize l = in.lower_rev;
    for (int k = 0; k < 4; ++k) {
      uint8_t c = (uint8_t)(4 - k);
      l += s;
      usize o = in.lower == 0 ? 0 : occ(in.lower - 1, c);
      s = occ(in.lower + in.size - 1, c) - o;
      out[k] = BiInterval{less[c] + o, l, s};
    }
  }

  // fmd_index.rs:77-91 (plain, non-transformed symbol)
  BiInterval backward_ext(const BiInterval& iv, uint8_t a) const {
    if (rank_of[a] < 0) return BiInterval{0, 0, 0};
    int r = rank_of[a];
    // The reference `.expect()`s here for ranks outside 1..=4 (e.g. 'X' or '$' in a read);
    // we return the empty interval instead of aborting.
    if (r < 1 || r > 4) return BiInterval{0, 0, 0};
    BiInterval out[4];
    extend_all(iv, out);
    return out[4 - r];
  }
  // fmd_index.rs:93-96
  BiInterval forward_ext(const BiInterval& iv, uint8_t a) const {
    return backward_ext(iv.swapped(), complement(a)).swapped();
  }
  uint8_t get_rev(uint8_t c) const { return back_transform[c]; }  // fmd_index.rs:103-105

  // SampledSuffixArray::get (index/mod.rs:160-187)
  bool sa_get(usize index, usize* out, Counters* ctr) const {
    if (index >= n) return false;
    if (!full_sa.empty() && sa_sample.empty()) { *out = full_sa[index]; return true; }
    usize pos = index, offset = 0;
    while (true) {
      if (pos % sa_rate == 0) {
        if (ctr) { ctr->lf_steps += 1; ctr->located += 1; }
        *out = sa_sample[pos / sa_rate] + offset;
        return true;
      }
      uint8_t c = bwt[pos];
      if (c == sentinel) {
        if (ctr) { ctr->lf_steps += 1; ctr->located += 1; }
        *out = extra_rows.at(pos) + offset;
        return true;
      }
      pos = less[c] + occ(pos - 1, c);
      offset += 1;
      if (ctr) ctr->lf_steps += 1;
    }
  }

  // FastaIdPositions::get_reference_identifier (index/mod.rs:55-75)
  bool reference_identifier(usize position, usize pattern_length, uint32_t* tid, uint64_t* rel) const {
    for (size_t i = 0; i < contigs.size(); ++i) {
      if (contigs[i].start <= position && position + pattern_length - 1 <= contigs[i].end) {
        *tid = (uint32_t)i;
        *rel = position - contigs[i].start;
        return true;
      }
    }
    return false;
  }
  // OriginalSymbols::get (index/mod.rs:206-209)
  bool original_symbol(usize idx, uint8_t* out) const {
    auto it = original_symbols.find(idx);
    if (it == original_symbols.end()) return false;
    *out = it->second;
    return true;
  }
};

// Suffix array by prefix doubling (own, simple algorithm; the product uses SA-IS).  Symbol order
// follows rust-bio's transform_text: the LAST sentinel is the smallest symbol, an earlier one the
// next smallest, then the ranks.
inline std::vector<usize> build_suffix_array(const std::vector<uint8_t>& text_ranks) {
  const usize n = text_ranks.size();
  std::vector<usize> sa(n), rnk(n), tmp(n);
  usize n_sent = 0;
  for (uint8_t c : text_ranks) n_sent += (c == 0);
  {
    usize s = n_sent;
    for (usize i = 0; i < n; ++i) {
      if (text_ranks[i] == 0) { s -= 1; rnk[i] = s; }
      else rnk[i] = text_ranks[i] + (n_sent - 1);
      sa[i] = i;
    }
  }
  for (usize h = 1;; h <<= 1) {
    auto key2 = [&](usize i) -> long long { return i + h < n ? (long long)rnk[i + h] : -1LL; };
    auto cmp = [&](usize a, usize b) {
      if (rnk[a] != rnk[b]) return rnk[a] < rnk[b];
      return key2(a) < key2(b);
    };
    std::sort(sa.begin(), sa.end(), cmp);
    tmp[sa[0]] = 0;
    for (usize i = 1; i < n; ++i) tmp[sa[i]] = tmp[sa[i - 1]] + (cmp(sa[i - 1], sa[i]) ? 1 : 0);
    rnk.swap(tmp);
    if (rnk[sa[n - 1]] == n - 1) break;
  }
  return sa;
}

struct GenomeInput {
  std::vector<std::string> names;
  std::vector<std::string> seqs;  // raw FASTA sequence per record (any case, IUPAC)
};

// IUPAC replacement callback: (symbol) -> base.  The reference draws from rand 0.9 StdRng
// (indexing.rs:79-93) which cannot be reproduced here; tests inject the draws they need.
typedef uint8_t (*ReplaceFn)(uint8_t symbol, void* user);

// indexing.rs:43-212 and utils.rs:12-33 share everything after the text has been assembled.
// `with_x`: production alphabet $ACGTX (indexing.rs:32) vs. test alphabet $ACGT (utils.rs).
inline void build_index(Index& ix, const GenomeInput& g, bool with_x, uint32_t occ_k, usize sa_rate,
                        bool keep_full_sa, ReplaceFn replace, void* replace_user) {
  std::vector<uint8_t> ref;
  for (const std::string& s : g.seqs)
    for (char ch : s) ref.push_back((uint8_t)std::toupper((unsigned char)ch));  // indexing.rs:60-66
  // run_apply (indexing.rs:217-256): runs >= 20 of a non-ACGT symbol -> 'X', shorter -> replaced
  ix.original_symbols.clear();
  {
    usize i = 0;
    while (i < ref.size()) {
      uint8_t sym = ref[i];
      usize run = 1;
      while (i + run < ref.size() && ref[i + run] == sym) ++run;
      bool in_alpha = sym == 'A' || sym == 'C' || sym == 'G' || sym == 'T';
      if (!in_alpha) {
        if (run < 20) {
          for (usize j = 0; j < run; ++j) {
            uint8_t rep = sym == 'U' ? (uint8_t)'T' : (replace ? replace(sym, replace_user) : (uint8_t)'A');
            ix.original_symbols[i + j] = sym;
            ref[i + j] = rep;
          }
        } else {
          for (usize j = 0; j < run; ++j) ref[i + j] = 'X';
        }
      }
      i += run;
    }
  }
  // contig map (indexing.rs:118-136)
  ix.contigs.clear();
  {
    uint64_t end = 0;
    for (size_t r = 0; r < g.seqs.size(); ++r) {
      end += g.seqs[r].size();
      ix.contigs.push_back(Contig{end - g.seqs[r].size(), end - 1, r < g.names.size() ? g.names[r] : std::string()});
    }
  }
  // text = fwd $ revcomp $  (indexing.rs:139-144, utils.rs:16-20)
  std::vector<uint8_t> text = ref;
  text.push_back('$');
  for (usize i = ref.size(); i-- > 0;) text.push_back(complement(ref[i]));
  text.push_back('$');
  // RankTransform over the sorted alphabet (indexing.rs:147-152)
  for (int& r : ix.rank_of) r = -1;
  std::string alpha = with_x ? "$ACGTX" : "$ACGT";
  std::sort(alpha.begin(), alpha.end());
  ix.back_transform.clear();
  for (size_t i = 0; i < alpha.size(); ++i) { ix.rank_of[(uint8_t)alpha[i]] = (int)i; ix.back_transform.push_back((uint8_t)alpha[i]); }
  for (uint8_t& c : text) c = (uint8_t)ix.rank_of[c];
  const usize n = text.size();
  ix.n = n;
  std::vector<usize> sa = build_suffix_array(text);
  // bwt (rust-bio bwt()): bwt[i] = text[sa[i]-1], wrapping
  ix.bwt.resize(n);
  for (usize i = 0; i < n; ++i) ix.bwt[i] = sa[i] > 0 ? text[sa[i] - 1] : text[n - 1];
  // sampled SA (index/mod.rs:91-128)
  ix.sa_rate = sa_rate;
  ix.sentinel = text[n - 1];
  ix.sa_sample.clear();
  ix.extra_rows.clear();
  for (usize i = 0; i < n; ++i) {
    if (i % sa_rate == 0) ix.sa_sample.push_back(sa[i]);
    else if (ix.bwt[i] == ix.sentinel) ix.extra_rows[i] = sa[i];
  }
  // less (rust-bio less()): size max_symbol + 2
  const usize m = alpha.size() - 1 + 2;
  ix.less.assign(m, 0);
  for (usize i = 0; i < n; ++i) ix.less[ix.bwt[i]] += 1;
  {
    usize acc = 0;
    for (usize c = 0; c < m; ++c) { usize v = ix.less[c]; ix.less[c] = acc; acc += v; }
  }
  // Occ::new(bwt, k, alphabet)
  ix.occ_k = occ_k;
  ix.occ_cp.assign(alpha.size(), std::vector<usize>());
  {
    std::vector<usize> cur(alpha.size(), 0);
    for (usize i = 0; i < n; ++i) {
      cur[ix.bwt[i]] += 1;
      if (i % occ_k == 0) for (size_t a = 0; a < alpha.size(); ++a) ix.occ_cp[a].push_back(cur[a]);
    }
  }
  // sentinel rows (fmd_index.rs:38-47)
  {
    int k = 0;
    ix.sentinel_occ[0] = ix.sentinel_occ[1] = 0;
    for (usize i = 0; i < n && k < 2; ++i) if (ix.bwt[i] == 0) ix.sentinel_occ[k++] = i;
  }
  if (keep_full_sa) ix.full_sa = sa; else ix.full_sa.clear();
}

// Rebuild the derived tables from externally supplied arrays (so the oracle can check the CUDA
// path on an index produced by the product's own builder at sizes its simple SA sort cannot reach).
inline void index_from_arrays(Index& ix, const uint8_t* bwt, usize n, bool with_x, uint32_t occ_k,
                              const uint64_t* sa_sample, usize n_samples, usize sa_rate,
                              const uint64_t* extra_rows_kv, usize n_extra) {
  ix.n = n;
  ix.bwt.assign(bwt, bwt + n);
  for (int& r : ix.rank_of) r = -1;
  std::string alpha = with_x ? "$ACGTX" : "$ACGT";
  ix.back_transform.clear();
  for (size_t i = 0; i < alpha.size(); ++i) { ix.rank_of[(uint8_t)alpha[i]] = (int)i; ix.back_transform.push_back((uint8_t)alpha[i]); }
  const usize m = alpha.size() - 1 + 2;
  ix.less.assign(m, 0);
  for (usize i = 0; i < n; ++i) ix.less[ix.bwt[i]] += 1;
  { usize acc = 0; for (usize c = 0; c < m; ++c) { usize v = ix.less[c]; ix.less[c] = acc; acc += v; } }
  ix.occ_k = occ_k;
  ix.occ_cp.assign(alpha.size(), std::vector<usize>());
  {
    std::vector<usize> cur(alpha.size(), 0);
    for (usize i = 0; i < n; ++i) {
      cur[ix.bwt[i]] += 1;
      if (i % occ_k == 0) for (size_t a = 0; a < alpha.size(); ++a) ix.occ_cp[a].push_back(cur[a]);
    }
  }
  { int k = 0; ix.sentinel_occ[0] = ix.sentinel_occ[1] = 0;
    for (usize i = 0; i < n && k < 2; ++i) if (ix.bwt[i] == 0) ix.sentinel_occ[k++] = i; }
  ix.sentinel = 0;
  ix.sa_rate = sa_rate;
  ix.sa_sample.assign(sa_sample, sa_sample + n_samples);
  ix.extra_rows.clear();
  for (usize i = 0; i < n_extra; ++i) ix.extra_rows[extra_rows_kv[2 * i]] = extra_rows_kv[2 * i + 1];
  ix.full_sa.clear();
}

// ------------------------------------------------------------------------------------------
// Sequence difference models (src/map/sequence_difference_models.rs)
// ------------------------------------------------------------------------------------------
enum SdmKind { SDM_SIMPLE = 0, SDM_VINDIJA = 1, SDM_TEST = 2 };
enum LibraryKind { LIB_SINGLE_STRANDED = 0, LIB_DOUBLE_STRANDED = 1 };

struct Sdm {
  int kind = SDM_SIMPLE;
  // SimpleAncientDnaModel (:104-114)
  int library = LIB_SINGLE_STRANDED;
  float five_prime_overhang = 0, three_prime_overhang = 0;  // DoubleStranded(x): both = x
  float ds_deamination_rate = 0, ss_deamination_rate = 0, divergence = 0;
  bool use_default_base_quality = false;
  float default_base_quality_prob = 0;
  float cache[256];
  // TestDifferenceModel (:396-401)
  float deam_score = 0, mm_score = 0, match_score = 0;
  // VindijaPwm (:340-346, :384-394)
  float ppm_ct[7] = {0.4f, 0.25f, 0.1f, 0.06f, 0.05f, 0.04f, 0.03f};
  float ppm_ct_default = 0.02f, subst_default = 0.0005f;

  static float qual2prob(uint8_t q) {  // :275-277
    return powf(10.0f, -(float)q / 10.0f) / 3.0f;
  }
  void init_simple(int lib, float f, float t, float d, float s, float div, bool ignore_q) {  // :279-333
    kind = SDM_SIMPLE; library = lib; five_prime_overhang = f; three_prime_overhang = t;
    ds_deamination_rate = d; ss_deamination_rate = s; divergence = div;
    use_default_base_quality = ignore_q;
    default_base_quality_prob = qual2prob(255);
    for (int q = 0; q < 256; ++q) cache[q] = qual2prob((uint8_t)q);
  }

  float get_simple(usize i, usize read_length, uint8_t from, uint8_t to, uint8_t q) const {  // :117-207
    const usize fp_dist = i, tp_dist = read_length - 1 - i;
    float seq_err = use_default_base_quality ? default_base_quality_prob : cache[q];
    float indep = fmaf(seq_err, -divergence, seq_err + divergence);
    float c_to_t = 0.f, g_to_a = 0.f;
    bool need_deam = (from == 'C' && (to == 'C' || to == 'T')) || (from == 'G' && (to == 'A' || to == 'G'));
    if (need_deam) {
      float p_fwd, p_rev;
      if (library == LIB_SINGLE_STRANDED) {
        float fpo = powi_f32(five_prime_overhang, (int)fp_dist + 1);
        float tpo = powi_f32(three_prime_overhang, (int)tp_dist + 1);
        p_fwd = fmaf(fpo, -tpo, fpo + tpo);
        p_rev = 0.0f;
      } else {
        p_fwd = powi_f32(five_prime_overhang, (int)fp_dist + 1);
        p_rev = powi_f32(five_prime_overhang, (int)tp_dist + 1);
      }
      c_to_t = fmaf(ss_deamination_rate, p_fwd, ds_deamination_rate * (1.0f - p_fwd));
      g_to_a = fmaf(ss_deamination_rate, p_rev, ds_deamination_rate * (1.0f - p_rev));
    }
    float v;
    switch (from) {
      case 'A': v = (to == 'A') ? fmaf(3.0f, -indep, 1.0f) : indep; break;
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
      case 'T': v = (to == 'T') ? fmaf(3.0f, -indep, 1.0f) : indep; break;
      default: v = indep;
    }
    return log2f(fmax_rs(v, std::numeric_limits<float>::epsilon()));
  }
  float get_vindija(usize i, usize read_length, uint8_t from, uint8_t to) const {  // :353-381
    float p;
    if (from == 'C') {
      usize k = std::min(i, read_length - (i + 1));
      float pct = k < 7 ? ppm_ct[k] : ppm_ct_default;
      if (to == 'T') p = pct; else if (to == 'C') p = 1.0f - pct; else p = subst_default;
    } else {
      p = (from == to) ? 1.0f - subst_default : subst_default;
    }
    return log2f(p);
  }
  float get_test(uint8_t from, uint8_t to) const {  // :409-419
    if (from == 'C' && to == 'T') return deam_score;
    if (from == to) return match_score;
    return mm_score;
  }
  float get(usize i, usize read_length, uint8_t from, uint8_t to, uint8_t q) const {
    switch (kind) {
      case SDM_SIMPLE: return get_simple(i, read_length, from, to, q);
      case SDM_VINDIJA: return get_vindija(i, read_length, from, to);
      default: return get_test(from, to);
    }
  }
  float representative_mismatch_penalty() const {  // :16-31
    return get(40, 80, 'T', 'A', 255) - get(40, 80, 'T', 'T', 255);
  }
  float get_min_penalty(usize i, usize read_length, uint8_t to, uint8_t q, bool only_mismatches) const {  // :34-57
    static const uint8_t ACGT[4] = {'A', 'C', 'G', 'T'};
    if (!only_mismatches) {
      if (!(to == 'A' || to == 'C' || to == 'G' || to == 'T')) return 0.0f;
    }
    float best = std::numeric_limits<float>::lowest();
    for (uint8_t base : ACGT) {
      if (only_mismatches && base == to) continue;
      best = fmax_rs(best, get(i, read_length, base, to, q));
    }
    return best;
  }
  int16_t find_alignment_start(usize pattern_length) const {  // :59-61 vs :209-211
    if (kind == SDM_SIMPLE) return (int16_t)pattern_length;
    return (int16_t)((int16_t)pattern_length / 2);
  }
};

// ------------------------------------------------------------------------------------------
// Mismatch bounds (src/map/mismatch_bounds.rs)
// ------------------------------------------------------------------------------------------
enum BoundKind { MB_CONTINUOUS = 0, MB_DISCRETE = 1, MB_TEST = 2 };

struct Bound {
  int kind = MB_DISCRETE;
  float repr_mm = 0;
  // Discrete (:123-128)
  float poisson_threshold = 0, base_error_rate = 0;
  float dcache[256];
  // Continuous (:77-82)
  float cutoff = 0, exponent = 0;
  float ccache[256];
  // TestBound (:264-267)
  float threshold = 0, representative_mm_bound = 0;

  static float calc_max_num_mismatches(usize read_length, float thr, float rate) {  // :209-236
    float lambda = (float)read_length * rate;
    float exp_minus_lambda = expf(-lambda);
    uint64_t last_k = 0;
    bool any = false;
    // k = 0 term
    float sum = exp_minus_lambda;
    if (1.0f - sum > thr) { last_k = 1; any = true; } else return 0.0f;
    float lambda_to_the_k = 1.0f;
    uint64_t k_factorial = 1;
    for (uint64_t k = 1; k <= (uint64_t)read_length; ++k) {
      lambda_to_the_k *= lambda;
      k_factorial *= k;  // wraps like release-mode Rust; take_while stops long before k = 21
      sum += lambda_to_the_k * exp_minus_lambda / (float)k_factorial;
      if (1.0f - sum > thr) last_k = k + 1; else break;
    }
    (void)any;
    return (float)last_k;
  }
  void init_discrete(float thr, float rate, float rmm) {  // :186-207
    kind = MB_DISCRETE; poisson_threshold = thr; base_error_rate = rate; repr_mm = rmm;
    for (usize i = 0; i < 256; ++i) dcache[i] = calc_max_num_mismatches(i + 17, thr, rate);
  }
  void init_continuous(float cut, float expo, float rmm) {  // :102-114
    kind = MB_CONTINUOUS; cutoff = cut; exponent = expo; repr_mm = rmm;
    for (usize i = 0; i < 256; ++i) ccache[i] = powf((float)i, expo);
  }
  void init_test(float thr, float rmb) { kind = MB_TEST; threshold = thr; representative_mm_bound = rmb; repr_mm = rmb; }

  float discrete_get(usize read_length) const {  // :238-255
    if (read_length < 17) return 0.0f;
    usize idx = read_length - 17;
    if (idx < 256) return dcache[idx];
    return calc_max_num_mismatches(read_length, poisson_threshold, base_error_rate);
  }
  float scale_read_length(usize read_length) const {  // :116-121
    if (read_length < 256) return ccache[read_length];
    return powf((float)read_length, exponent);
  }
  bool reject(float value, usize read_length) const {
    switch (kind) {
      case MB_CONTINUOUS: return (value / scale_read_length(read_length)) < cutoff;        // :85-87
      case MB_DISCRETE: return value < discrete_get(read_length) * repr_mm;                 // :131-134
      default: return value < threshold;                                                    // :270-272
    }
  }
  bool reject_iterative(float value, float reference) const {
    if (kind == MB_TEST) return false;          // :274-276
    return value < reference + repr_mm;          // :89-91, :136-138
  }
  float remaining_frac_of_repr_mm(float value, usize read_length) const {
    switch (kind) {
      case MB_CONTINUOUS: {                      // :93-97
        float s = scale_read_length(read_length);
        return (cutoff - value / s) / (repr_mm / s);
      }
      case MB_DISCRETE:                          // :140-144
        return fmaf(discrete_get(read_length), repr_mm, -value) / repr_mm;
      default:                                   // :278-280
        return (threshold - value) / representative_mm_bound;
    }
  }
};

struct Params {  // AlignmentParameters (src/map/mod.rs:21-31)
  Sdm sdm;
  Bound mb;
  float penalty_gap_open = 0, penalty_gap_extend = 0;
  uint8_t gap_dist_ends = 5, max_num_gaps_open = 2;
  bool stack_limit_abort = false;
  uint32_t stack_limit = 2000000;       // mapping.rs:53
  uint32_t edit_tree_limit = 10000000;  // mapping.rs:54
};

// ------------------------------------------------------------------------------------------
// Heaps
// ------------------------------------------------------------------------------------------
// min_max_heap::MinMaxHeap; T needs `float key() const`.
template <class T>
struct MinMaxHeap {
  std::vector<T> d;
  static bool lt(const T& a, const T& b) { return a.key() < b.key(); }
  static bool gt(const T& a, const T& b) { return a.key() > b.key(); }
  static bool on_min_level(size_t i) {  // level = floor(log2(i+1)); even = min level
    int level = 63 - __builtin_clzll((unsigned long long)(i + 1));
    return (level & 1) == 0;
  }
  size_t len() const { return d.size(); }
  void clear() { d.clear(); }
  void push(const T& x) {
    d.push_back(x);
    size_t i = d.size() - 1;
    T e = d[i];
    bool min_level = on_min_level(i);
    bool climb_max;  // which comparator the grandparent climb uses
    if (i > 0) {
      size_t p = (i - 1) / 2;
      if (min_level) {
        if (gt(e, d[p])) { d[i] = d[p]; i = p; climb_max = true; } else climb_max = false;
      } else {
        if (lt(e, d[p])) { d[i] = d[p]; i = p; climb_max = false; } else climb_max = true;
      }
    } else {
      climb_max = !min_level;
    }
    while (i >= 3) {
      size_t gp = ((i - 1) / 2 - 1) / 2;
      if (climb_max ? gt(e, d[gp]) : lt(e, d[gp])) { d[i] = d[gp]; i = gp; } else break;
    }
    d[i] = e;
  }
  template <bool MAX>
  void trickle_down(size_t i) {
    auto better = [](const T& a, const T& b) { return MAX ? gt(a, b) : lt(a, b); };
    const size_t n = d.size();
    T e = d[i];
    while (true) {
      size_t best = (size_t)-1;
      const T* bk = &e;
      const size_t cand[6] = {2 * i + 1, 2 * i + 2, 4 * i + 3, 4 * i + 4, 4 * i + 5, 4 * i + 6};
      for (int c = 0; c < 6; ++c) {
        if (cand[c] >= n) break;
        if (better(d[cand[c]], *bk)) { best = cand[c]; bk = &d[cand[c]]; }
      }
      if (best == (size_t)-1) break;
      bool was_child = best <= 2 * i + 2;
      d[i] = d[best];
      i = best;
      if (was_child) break;
      size_t p = (i - 1) / 2;
      if (better(d[p], e)) std::swap(e, d[p]);
    }
    d[i] = e;
  }
  bool pop_max(T* out) {
    size_t n = d.size();
    if (n == 0) return false;
    size_t m = n == 1 ? 0 : (n == 2 ? 1 : (gt(d[1], d[2]) ? 1 : 2));
    T item = d.back();
    d.pop_back();
    if (m < d.size()) { std::swap(item, d[m]); trickle_down<true>(m); }
    *out = item;
    return true;
  }
  bool pop_min(T* out) {
    if (d.empty()) return false;
    T item = d.back();
    d.pop_back();
    if (!d.empty()) { std::swap(item, d[0]); trickle_down<false>(0); }
    *out = item;
    return true;
  }
};

// std::collections::BinaryHeap (max-heap); T needs `float key() const`.
template <class T>
struct BinaryHeap {
  std::vector<T> d;
  size_t len() const { return d.size(); }
  bool empty() const { return d.empty(); }
  const T* peek() const { return d.empty() ? nullptr : &d[0]; }
  void sift_up(size_t start, size_t pos) {
    T e = d[pos];
    while (pos > start) {
      size_t parent = (pos - 1) / 2;
      if (e.key() <= d[parent].key()) break;
      d[pos] = d[parent];
      pos = parent;
    }
    d[pos] = e;
  }
  void push(const T& x) { d.push_back(x); sift_up(0, d.size() - 1); }
  bool pop(T* out) {
    if (d.empty()) return false;
    T item = d.back();
    d.pop_back();
    if (!d.empty()) {
      std::swap(item, d[0]);
      // sift_down_to_bottom(0)
      size_t end = d.size(), pos = 0;
      T e = d[0];
      size_t child = 1;
      while (child + 2 <= end) {  // child <= end.saturating_sub(2)
        if (d[child].key() <= d[child + 1].key()) child += 1;
        d[pos] = d[child];
        pos = child;
        child = 2 * pos + 1;
      }
      if (child + 1 == end) { d[pos] = d[child]; pos = child; }
      d[pos] = e;
      sift_up(0, pos);
    }
    *out = item;
    return true;
  }
  // into_sorted_vec: ascending; the caller pops from the back (mapping.rs:419-421)
  std::vector<T> into_sorted_vec() {
    size_t end = d.size();
    while (end > 1) {
      end -= 1;
      std::swap(d[0], d[end]);
      // sift_down_range(0, end)
      size_t pos = 0;
      T e = d[0];
      size_t child = 1;
      bool done = false;
      while (child + 2 <= end) {
        if (d[child].key() <= d[child + 1].key()) child += 1;
        if (e.key() >= d[child].key()) { done = true; break; }
        d[pos] = d[child];
        pos = child;
        child = 2 * pos + 1;
      }
      if (!done && child + 1 == end && e.key() < d[child].key()) { d[pos] = d[child]; pos = child; }
      d[pos] = e;
    }
    std::vector<T> out;
    out.swap(d);
    return out;
  }
};

// ------------------------------------------------------------------------------------------
// Edit operations, backtrack tree (record.rs:226-231, backtrack_tree.rs, slab)
// ------------------------------------------------------------------------------------------
enum EditKind : uint8_t { ED_INSERTION = 0, ED_DELETION = 1, ED_MATCH = 2, ED_MISMATCH = 3 };
struct EditOp {
  uint16_t pos = 0;
  uint8_t kind = ED_MATCH;
  uint8_t base = 0;  // reference base for Deletion / Mismatch
};

struct Tree {
  struct Node { EditOp value; uint32_t parent; bool occupied; uint32_t next_free; };
  std::vector<Node> entries;
  uint32_t next = 0;   // slab free-list head
  uint32_t count = 0;  // slab len
  uint32_t insert(const EditOp& v, uint32_t parent) {  // slab::insert
    uint32_t key = next;
    if (key == entries.size()) {
      entries.push_back(Node{v, parent, true, 0});
      next = key + 1;
    } else {
      next = entries[key].next_free;
      entries[key] = Node{v, parent, true, 0};
    }
    count += 1;
    return key;
  }
  void remove(uint32_t key) {  // backtrack_tree.rs:49-53
    if (key != 0) {
      entries[key].occupied = false;
      entries[key].next_free = next;
      next = key;
      count -= 1;
    }
  }
  uint32_t clear() {  // backtrack_tree.rs:93-97
    entries.clear(); next = 0; count = 0;
    return insert(EditOp(), 0);
  }
  uint32_t add_node(const EditOp& v, uint32_t parent) { return insert(v, parent); }
  uint32_t len() const { return count; }
};

// extract_edit_operations (record.rs:465-500)
inline std::vector<EditOp> extract_edit_operations(uint32_t end_node, const Tree& tree, int16_t alignment_start) {
  std::map<uint16_t, std::vector<EditOp>> buckets;
  uint32_t state = end_node;
  while (state != 0) {
    const Tree::Node& node = tree.entries[state];
    if (!node.occupied) break;  // slab.get() -> None
    buckets[node.value.pos].push_back(node.value);
    state = node.parent;
  }
  std::vector<EditOp> out;
  for (auto& kv : buckets) {
    if (kv.first < (uint16_t)alignment_start) out.insert(out.end(), kv.second.begin(), kv.second.end());
    else out.insert(out.end(), kv.second.rbegin(), kv.second.rend());
  }
  return out;
}

inline usize effective_len(const std::vector<EditOp>& ops) {  // record.rs:269-279
  usize n = 0;
  for (const EditOp& o : ops) n += (o.kind != ED_INSERTION);
  return n;
}
inline usize read_len(const std::vector<EditOp>& ops) {  // record.rs:430-440
  usize n = 0;
  for (const EditOp& o : ops) n += (o.kind != ED_DELETION);
  return n;
}

struct CigarOp { uint8_t kind; uint32_t len; };  // kind: 'M','I','D'

// EditOperationsTrack::to_bam_fields (record.rs:282-428)
inline void to_bam_fields(const std::vector<EditOp>& track_in, bool backward_strand, usize absolute_pos,
                          const Index& ix, std::vector<CigarOp>* cigar, std::string* md, uint16_t* nm) {
  uint32_t num_matches = 0, num_operations = 1;
  uint16_t edit_distance = 0;
  bool have_last = false;
  EditOp last;
  cigar->clear();
  md->clear();
  auto cigar_kind = [](const EditOp& o) -> uint8_t {
    return o.kind == ED_INSERTION ? 'I' : (o.kind == ED_DELETION ? 'D' : 'M');
  };
  auto comp_if = [&](uint8_t b) { return backward_strand ? complement(b) : b; };
  auto add_md = [&](const EditOp* op, const EditOp* lop, uint32_t k) -> uint32_t {
    if (!op) { *md += std::to_string(k); return k; }
    switch (op->kind) {
      case ED_MATCH: k += 1; break;
      case ED_MISMATCH: *md += std::to_string(k); md->push_back((char)comp_if(op->base)); k = 0; break;
      case ED_INSERTION: break;
      case ED_DELETION:
        if (lop && lop->kind == ED_DELETION) md->push_back((char)comp_if(op->base));
        else { *md += std::to_string(k); md->push_back('^'); md->push_back((char)comp_if(op->base)); }
        k = 0;
        break;
    }
    return k;
  };
  const size_t n = track_in.size();
  for (size_t i = 0; i < n; ++i) {
    EditOp op = backward_strand ? track_in[n - 1 - i] : track_in[i];
    uint8_t orig;
    switch (op.kind) {
      case ED_INSERTION: break;
      case ED_MATCH:
        if (ix.original_symbol(absolute_pos + i, &orig)) { op.kind = ED_MISMATCH; op.base = orig; }
        break;
      case ED_DELETION:
      case ED_MISMATCH:
        if (ix.original_symbol(absolute_pos + i, &orig)) op.base = orig;
        break;
    }
    if (op.kind != ED_MATCH) edit_distance += 1;
    num_matches = add_md(&op, have_last ? &last : nullptr, num_matches);
    if (have_last) {
      bool same_class;
      switch (op.kind) {
        case ED_MATCH: case ED_MISMATCH: same_class = (last.kind == ED_MATCH || last.kind == ED_MISMATCH); break;
        case ED_INSERTION: same_class = last.kind == ED_INSERTION; break;
        default: same_class = last.kind == ED_DELETION; break;
      }
      if (same_class) num_operations += 1;
      else { cigar->push_back(CigarOp{cigar_kind(last), num_operations}); num_operations = 1; last = op; }
    } else {
      last = op;
      have_last = true;
    }
  }
  if (have_last) cigar->push_back(CigarOp{cigar_kind(last), num_operations});
  add_md(nullptr, nullptr, num_matches);
  *nm = edit_distance;
}

// ------------------------------------------------------------------------------------------
// D array (src/map/bi_d_array.rs)
// ------------------------------------------------------------------------------------------
struct BiDArray {
  std::vector<float> d_composite;
  usize split = 0;

  // compute_part (:104-198), materialised to `want` elements.
  static std::vector<float> compute_part(const uint8_t* part, const uint8_t* quals, usize part_len, bool dir_forward,
                                         usize full_len, uint16_t initial_skip, const Params& p, const Index& ix,
                                         usize want, Counters* ctr) {
    std::vector<float> out;
    out.reserve(want);
    for (usize i = 0; i < (usize)initial_skip + 1 && out.size() < want; ++i) out.push_back(0.0f);
    float z = 0.0f;
    int16_t last_mismatch_pos = (int16_t)initial_skip - 1;
    BiInterval interval = ix.init_interval();
    for (usize index = initial_skip; index < part_len && out.size() < want; ++index) {
      uint8_t base = dir_forward ? part[index] : part[part_len - 1 - index];
      interval = dir_forward ? ix.forward_ext(interval, base) : ix.backward_ext(interval, base);
      if (ctr) ctr->d_ext_steps += 1;
      if (interval.size < 1) {
        float m = std::numeric_limits<float>::lowest();
        for (usize j = (usize)(last_mismatch_pos + 1); j <= index; ++j) {
          uint8_t base_j = dir_forward ? part[j] : part[part_len - 1 - j];
          uint8_t qual_j = dir_forward ? quals[j] : quals[part_len - 1 - j];
          usize idx = dir_forward ? j : full_len - 1 - j;  // directed_index(j, full_len, dir)
          float best_mm = p.sdm.get_min_penalty(idx, full_len, base_j, qual_j, true);
          float optimal = p.sdm.get_min_penalty(idx, full_len, base_j, qual_j, false);
          float mm_retval = best_mm - optimal;
          float v = (std::min(idx, full_len - idx - 1) >= (usize)p.gap_dist_ends)
                        ? fmax_rs(mm_retval, p.penalty_gap_extend) : mm_retval;
          m = fmax_rs(m, v);
        }
        z += m;
        interval = ix.init_interval();
        last_mismatch_pos = (int16_t)index;
      }
      out.push_back(z);
    }
    return out;
  }

  void build(const uint8_t* pattern, const uint8_t* quals, usize len, usize split_, const Params& p, const Index& ix,
             Counters* ctr) {  // :24-99
    split = split_;
    d_composite.assign(len, 0.0f);
    const int MAX_OFFSET = 15;
    for (usize pos = 0; pos < len; ++pos) d_composite[pos] = 0.0f;
    for (int off = 0; off < MAX_OFFSET; ++off) {
      std::vector<float> part = compute_part(pattern, quals, split, true, len, (uint16_t)off, p, ix, split, ctr);
      for (usize i = 0; i < split; ++i) d_composite[i] = fmin_rs(d_composite[i], part[i]);
    }
    for (int off = 0; off < MAX_OFFSET; ++off) {
      std::vector<float> part =
          compute_part(pattern + split, quals + split, len - split, false, len, (uint16_t)off, p, ix, len - split, ctr);
      for (usize i = 0; i < len - split; ++i) d_composite[split + i] = fmin_rs(d_composite[split + i], part[i]);
    }
  }

  float get(int16_t backward_index, int16_t forward_index) const {  // :200-224
    float d_rev = 0.0f, d_fwd = 0.0f;
    if (backward_index >= 0 && (usize)backward_index < d_composite.size()) d_rev = d_composite[backward_index];
    usize need = 1 + (usize)forward_index;  // `forward_index as usize`
    if (forward_index >= 0 && d_composite.size() >= need) {
      usize idx = d_composite.size() - need + split;
      if (idx < d_composite.size()) d_fwd = d_composite[idx];
    }
    return d_rev + d_fwd;
  }
};

// ------------------------------------------------------------------------------------------
// k_mismatch_search (src/map/mapping.rs:932-1383, src/map/mod.rs:33-137)
// ------------------------------------------------------------------------------------------
enum GapState : uint8_t { GAP_INSERTION = 0, GAP_DELETION = 1, GAP_CLOSED = 2 };

struct Frame {  // MismatchSearchStackFrame (mod.rs:105-115)
  BiInterval current_interval;
  int16_t start = 0, len = 0;
  GapState gap_forwards = GAP_CLOSED, gap_backwards = GAP_CLOSED;
  uint8_t num_gaps_open = 0;
  float alignment_score = 0;
  uint32_t edit_node_id = 0;
  float key() const { return alignment_score; }
};

struct Hit {  // HitInterval (mod.rs:34-39)
  BiInterval interval;
  float alignment_score = 0;
  std::vector<EditOp> edit_operations;
  float key() const { return alignment_score; }
};

struct Scratch {
  MinMaxHeap<Frame> stack;
  Tree tree;
};

inline void check_and_push(Frame f, usize pattern_len, int16_t alignment_start_pos, const EditOp& op, Scratch& s,
                           BinaryHeap<Hit>& hits, const Params& p) {  // mapping.rs:932-987
  if (const Hit* best = hits.peek()) {
    if (p.mb.reject_iterative(f.alignment_score, best->alignment_score)) return;
  }
  if (f.num_gaps_open > p.max_num_gaps_open) return;
  f.edit_node_id = s.tree.add_node(op, f.edit_node_id);
  if ((usize)f.len == pattern_len) {
    Hit h;
    h.interval = f.current_interval;
    h.alignment_score = f.alignment_score;
    h.edit_operations = extract_edit_operations(f.edit_node_id, s.tree, alignment_start_pos);
    hits.push(h);
    return;
  }
  s.stack.push(f);
}

inline BinaryHeap<Hit> k_mismatch_search(const uint8_t* pattern, const uint8_t* quals, usize L, const Params& p,
                                         const Index& ix, Scratch& s, Counters* ctr) {
  static const uint8_t TGCA[4] = {'T', 'G', 'C', 'A'};  // b"ACGT".iter().rev()
  // An empty read makes the reference index optimal_penalties[0] out of bounds (panic="abort");
  // both the oracle and the CUDA path report such a read as unmapped instead.
  if (L == 0) return BinaryHeap<Hit>();
  const int16_t alignment_start_pos = p.sdm.find_alignment_start(L);
  BiDArray bi_d;
  bi_d.build(pattern, quals, L, (usize)alignment_start_pos, p, ix, ctr);
  std::vector<float> optimal(L);  // compute_optimal_scores (:572-588)
  for (usize i = 0; i < L; ++i) optimal[i] = p.sdm.get_min_penalty(i, L, pattern[i], quals[i], false);
  BinaryHeap<Hit> hits;
  s.stack.clear();
  uint32_t root = s.tree.clear();
  {
    Frame f;
    f.current_interval = ix.init_interval();
    f.start = alignment_start_pos; f.len = 0;
    f.alignment_score = 0.0f; f.edit_node_id = root;
    s.stack.push(f);
  }
  float mm_scores[4] = {0, 0, 0, 0};
  Frame sf;
  while (s.stack.pop_max(&sf)) {
    if (ctr) ctr->frames_popped += 1;
    int16_t j, d_k, d_l;
    bool forward;
    if (sf.start <= (int16_t)L - sf.start - sf.len) {  // :1077-1097
      j = sf.start + sf.len; forward = true; d_k = sf.start; d_l = sf.start + sf.len;
    } else {
      j = sf.start - 1; forward = false; d_k = sf.start - 1; d_l = sf.start + sf.len - 1;
    }
    const float optimal_penalty = optimal[j];
    BiInterval ext_interval;
    GapState ins_b, ins_f, del_b, del_f, cl_b, cl_f;
    float insertion_score, deletion_score;
    uint8_t num_gaps_open;
    if (forward) {  // :1116-1153
      ext_interval = sf.current_interval.swapped();
      ins_b = sf.gap_backwards; ins_f = GAP_INSERTION;
      del_b = sf.gap_backwards; del_f = GAP_DELETION;
      cl_b = sf.gap_backwards; cl_f = GAP_CLOSED;
      insertion_score = (sf.gap_forwards == GAP_INSERTION ? p.penalty_gap_extend
                                                          : p.penalty_gap_open + p.penalty_gap_extend) + sf.alignment_score;
      deletion_score = (sf.gap_forwards == GAP_DELETION ? p.penalty_gap_extend
                                                        : p.penalty_gap_open + p.penalty_gap_extend) + sf.alignment_score;
      for (int k = 0; k < 4; ++k)
        mm_scores[k] = p.sdm.get(j, L, complement(TGCA[k]), pattern[j], quals[j]) - optimal_penalty + sf.alignment_score;
      num_gaps_open = sf.gap_forwards == GAP_CLOSED ? sf.num_gaps_open + 1 : sf.num_gaps_open;
    } else {  // :1154-1191
      ext_interval = sf.current_interval;
      ins_b = GAP_INSERTION; ins_f = sf.gap_forwards;
      del_b = GAP_DELETION; del_f = sf.gap_forwards;
      cl_b = GAP_CLOSED; cl_f = sf.gap_forwards;
      insertion_score = (sf.gap_backwards == GAP_INSERTION ? p.penalty_gap_extend
                                                           : p.penalty_gap_open + p.penalty_gap_extend) + sf.alignment_score;
      deletion_score = (sf.gap_backwards == GAP_DELETION ? p.penalty_gap_extend
                                                         : p.penalty_gap_open + p.penalty_gap_extend) + sf.alignment_score;
      for (int k = 0; k < 4; ++k)
        mm_scores[k] = p.sdm.get(j, L, TGCA[k], pattern[j], quals[j]) - optimal_penalty + sf.alignment_score;
      num_gaps_open = sf.gap_backwards == GAP_CLOSED ? sf.num_gaps_open + 1 : sf.num_gaps_open;
    }
    const float lower_bound = bi_d.get(d_k, d_l);  // :1195
    if (const Hit* best = hits.peek()) {           // :1201-1208
      if (p.mb.reject_iterative(sf.alignment_score + lower_bound, best->alignment_score)) break;
    }
    // insertion (:1213-1242)
    if (!p.mb.reject(insertion_score + lower_bound, L) &&
        std::min<int16_t>(j, (int16_t)L - j - 1) >= (int16_t)p.gap_dist_ends) {
      Frame c = sf;
      c.start = forward ? sf.start : sf.start - 1;
      c.len = sf.len + 1;
      c.gap_backwards = ins_b; c.gap_forwards = ins_f;
      c.alignment_score = insertion_score;
      c.num_gaps_open = num_gaps_open;
      check_and_push(c, L, alignment_start_pos, EditOp{(uint16_t)j, ED_INSERTION, 0}, s, hits, p);
    }
    // extension (:1245-1339)
    BiInterval ext[4];
    ix.extend_all(ext_interval, ext);
    for (int k = 0; k < 4; ++k) {
      BiInterval ip = ext[k];
      if (ip.size < 1) continue;
      uint8_t c;
      if (forward) { ip = ip.swapped(); c = complement(ix.get_rev((uint8_t)(4 - k))); }
      else c = ix.get_rev((uint8_t)(4 - k));
      {
        int16_t dist_5 = forward ? j : j + 1;
        int16_t dist_3 = (int16_t)L - dist_5;
        int16_t dist = std::min(dist_5, dist_3);
        if (!p.mb.reject(deletion_score + lower_bound, L) && dist >= (int16_t)p.gap_dist_ends) {
          Frame ch = sf;
          ch.current_interval = ip;
          ch.gap_backwards = del_b; ch.gap_forwards = del_f;
          ch.alignment_score = deletion_score;
          ch.num_gaps_open = num_gaps_open;
          check_and_push(ch, L, alignment_start_pos, EditOp{(uint16_t)j, ED_DELETION, c}, s, hits, p);
        }
      }
      if (!p.mb.reject(mm_scores[k] + lower_bound, L)) {
        Frame ch = sf;
        ch.current_interval = ip;
        ch.start = forward ? sf.start : sf.start - 1;
        ch.len = sf.len + 1;
        ch.gap_backwards = cl_b; ch.gap_forwards = cl_f;
        ch.alignment_score = mm_scores[k];
        EditOp op = (c == pattern[j]) ? EditOp{(uint16_t)j, ED_MATCH, 0} : EditOp{(uint16_t)j, ED_MISMATCH, c};
        check_and_push(ch, L, alignment_start_pos, op, s, hits, p);
      }
    }
    if (ctr) { ctr->max_stack = std::max<uint64_t>(ctr->max_stack, s.stack.len()); }
    // early exits (:1348-1355)
    if (hits.len() > 9 || (hits.peek() && hits.peek()->interval.size > 1)) break;
    // limits (:1358-1380)
    if (s.stack.len() > p.stack_limit || s.tree.len() > p.edit_tree_limit) {
      if (ctr) ctr->limit_hit += 1;
      if (p.stack_limit_abort) break;
      long long excess = std::max((long long)s.stack.len() - (long long)p.stack_limit,
                                  (long long)s.tree.len() - (long long)p.edit_tree_limit);
      for (long long e = 0; e < excess; ++e) {
        Frame mn;
        if (s.stack.pop_min(&mn)) s.tree.remove(mn.edit_node_id);
      }
    }
  }
  if (ctr) ctr->tree_nodes += s.tree.len();
  return hits;
}

// ------------------------------------------------------------------------------------------
// PrRange (src/map/prrange.rs)
// ------------------------------------------------------------------------------------------
struct PrRange {
  usize start = 0, l = 0, m = 0, a = 0, x = 0, seed = 0, count = 0;
  bool valid = false;

  static bool is_prime(usize n) {
    if (n <= 1) return false;
    if (n <= 3) return true;
    if (n % 2 == 0 || n % 3 == 0) return false;
    for (usize i = 5; i * i <= n; i += 6)
      if (n % i == 0 || n % (i + 2) == 0) return false;
    return true;
  }
  static usize next_prime(usize n) {
    usize p = n + 1;
    if (p <= 2) return 2;
    if (p % 2 == 0) p += 1;
    while (!is_prime(p)) p += 2;
    return p;
  }
  static bool checked_pow_mod(usize base, usize exponent, usize modulus, usize* out) {
    if (modulus == 1) { *out = 0; return true; }
    unsigned __int128 sq = (unsigned __int128)(modulus - 1) * (modulus - 1);
    if (sq >> 64) return false;
    usize result = 1;
    base %= modulus;
    while (exponent > 0) {
      if (exponent % 2 == 1) result = (result * base) % modulus;
      exponent >>= 1;
      base = (base * base) % modulus;
    }
    *out = result;
    return true;
  }
  // PrimeFactorIterator (:118-159): distinct prime factors, as that iterator yields them
  static std::vector<usize> prime_factors(usize n0) {
    std::vector<usize> out;
    usize n = n0, i = 2, step = 1, last = 0;
    while (true) {
      if (n <= 3) return out;
      bool yielded = false;
      while (i * i <= n && !yielded) {
        while (n > 1 && !yielded) {
          while (n % i == 0) {
            if (i > last) { out.push_back(i); last = i; yielded = true; break; }
            n /= i;
          }
          if (yielded) break;
          i += step;
          step = 2;
        }
      }
      if (!yielded) return out;
    }
  }
  static bool is_primitive_root(usize a, usize n, bool* ok) {
    usize phi = n - 1;
    for (usize pf : prime_factors(phi)) {
      usize r;
      if (!checked_pow_mod(a, phi / pf, n, &r)) { *ok = false; return false; }
      if (r == 1) { *ok = true; return false; }
    }
    *ok = true;
    return true;
  }
  static PrRange try_new(usize start, usize end, usize seed) {  // :41-70
    PrRange r;
    usize l = end > start ? end - start : 0;
    if (l == 0) return r;
    usize m = next_prime(l);
    usize a = 2;
    while (true) {
      bool ok;
      bool pr = is_primitive_root(a, m, &ok);
      if (!ok) return r;
      if (pr) break;
      a += 1;
    }
    usize sd = std::max<usize>(seed % l, 1);
    r.start = start; r.l = l; r.m = m; r.a = a; r.x = sd; r.seed = sd; r.count = 0; r.valid = true;
    return r;
  }
  bool next(usize* out) {  // :19-37
    if (count == 0 && l == 1) { count += 1; *out = start; return true; }
    while (true) {
      usize prev_x = x;
      x = (a * x) % m;
      if (count > 0 && prev_x == seed) return false;
      if (prev_x <= l) { count += 1; *out = prev_x - 1 + start; return true; }
    }
  }
};

// ------------------------------------------------------------------------------------------
// Epilogue (src/map/mapping.rs:402-718)
// ------------------------------------------------------------------------------------------
// The reference draws `rng.next_u32()` from an UNSEEDED thread-local generator (mapping.rs:273,605).
// For reproducible parity both the oracle and the CUDA path derive the k-th draw of a read from a
// caller-supplied per-read seed with this mixer (splitmix64 finaliser).
inline uint32_t draw_u32(uint32_t read_seed, uint32_t k) {
  uint64_t z = ((uint64_t)read_seed << 32 | k) + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (uint32_t)(z >> 32);
}

struct Coord {  // IntToCoordOutput (mod.rs:155-164)
  uint32_t tid = 0;
  uint64_t relative_pos = 0;
  usize absolute_pos = 0;
  bool backward = false;
  usize num_skipped = 0;
  const Hit* interval = nullptr;
};

// Lazy iterator over the valid coordinates of one hit interval (mapping.rs:590-649)
struct CoordIter {
  const Hit* hit = nullptr;
  const Index* ix = nullptr;
  PrRange pr;
  usize enum_i = 0;
  usize eff_len = 0;
  Counters* ctr = nullptr;
  bool next(Coord* out) {
    usize sar_pos;
    while (pr.valid && pr.next(&sar_pos)) {
      usize i = enum_i++;
      usize abs;
      if (!ix->sa_get(sar_pos, &abs, ctr)) continue;
      const usize strand_len = ix->n / 2;
      bool backward = false;
      if (abs >= strand_len) { abs = ix->n - abs - eff_len - 1; backward = true; }
      uint32_t tid; uint64_t rel;
      if (ix->reference_identifier(abs, eff_len, &tid, &rel)) {
        out->tid = tid; out->relative_pos = rel; out->absolute_pos = abs; out->backward = backward;
        out->num_skipped = i; out->interval = hit;
        return true;
      }
    }
    return false;
  }
};
inline CoordIter interval2coordinate(const Hit& h, const Index& ix, uint32_t rng_draw, Counters* ctr) {
  CoordIter it;
  it.hit = &h; it.ix = &ix; it.ctr = ctr;
  it.eff_len = effective_len(h.edit_operations);
  it.pr = PrRange::try_new(h.interval.lower, h.interval.lower + h.interval.size, (usize)rng_draw);
  return it;
}
inline bool interval_cross_check(const BiInterval& a, const BiInterval& b) {  // :651-653
  return a.size == b.size && (a.lower == b.lower || a.lower_rev == b.lower_rev);
}

inline uint8_t estimate_mapping_quality(const Hit& best, usize best_size, const std::vector<Hit>& others,
                                        const Params& p) {  // :658-718
  const uint8_t MAX_MAPQ = 37, MIN_MAPQ_UNIQ = 20;
  float prob_best = exp2f(best.alignment_score);
  float ap;
  if (best_size > 1) ap = 1.0f / (float)best_size;
  else {
    float acc = 0.0f;
    for (const Hit& o : others) {
      if (interval_cross_check(best.interval, o.interval)) continue;
      acc = fmaf(exp2f(o.alignment_score), (float)o.interval.size, acc);
    }
    ap = prob_best / (prob_best + acc);
  }
  // f32::clamp(0,1): NaN stays NaN
  if (ap < 0.0f) ap = 0.0f;
  if (ap > 1.0f) ap = 1.0f;
  float q = -10.0f * log10f(1.0f - ap);
  q = fmin_rs(q, (float)MAX_MAPQ);
  float rq = roundf(q);
  uint8_t mq = std::isnan(rq) ? 0 : (rq <= 0.0f ? 0 : (rq >= 255.0f ? 255 : (uint8_t)rq));  // `as u8` saturates
  if (mq == MAX_MAPQ) {
    float frac = fmin_rs(p.mb.remaining_frac_of_repr_mm(best.alignment_score, read_len(best.edit_operations)), 1.0f);
    float scaled = fmaf((float)(MAX_MAPQ - MIN_MAPQ_UNIQ), frac, (float)MIN_MAPQ_UNIQ);
    float rs = roundf(scaled);
    return std::isnan(rs) ? 0 : (rs <= 0.0f ? 0 : (rs >= 255.0f ? 255 : (uint8_t)rs));
  }
  return mq;
}

struct AltHit {
  uint32_t tid; uint64_t relative_pos; usize absolute_pos; bool backward;
  std::vector<CigarOp> cigar; std::string md; uint16_t nm; usize interval_size; float score;
};

struct Record {  // the fields of the BAM record that the hot path decides (mapping.rs:522-566,722-927)
  bool mapped = false;
  uint32_t tid = 0;
  uint64_t pos = 0;       // 0-based relative position
  usize absolute_pos = 0;
  bool backward = false;
  uint8_t mapq = 0;
  float alignment_score = 0;
  std::vector<CigarOp> cigar;
  std::string md;
  uint16_t nm = 0;
  int32_t x0 = 0, x1 = 0;
  float xs = 0;
  char xt = 'N';
  std::string xa;
  std::vector<AltHit> alts;
  BiInterval best_interval;
};

inline std::string cigar_string(const std::vector<CigarOp>& c) {
  std::string s;
  for (const CigarOp& o : c) { s += std::to_string(o.len); s.push_back((char)o.kind); }
  return s;
}
inline std::string format_score_2(float v) {  // Rust `{:.2}`
  char buf[64];
  snprintf(buf, sizeof buf, "%.2f", (double)v);
  return buf;
}

inline Record intervals_to_record(BinaryHeap<Hit> heap, const Index& ix, const Params& p, uint32_t read_seed,
                                  Counters* ctr) {  // mapping.rs:402-567
  Record rec;
  uint32_t draw_k = 0;
  std::vector<Hit> intervals = heap.into_sorted_vec();
  while (!intervals.empty()) {
    Hit best = std::move(intervals.back());
    intervals.pop_back();
    CoordIter best_it = interval2coordinate(best, ix, draw_u32(read_seed, draw_k++), ctr);
    Coord bc;
    if (!best_it.next(&bc)) continue;  // :541-544
    usize updated_size = best.interval.size - bc.num_skipped;
    // XA (:436-491): remaining positions of the best hit, then sub-optimal hits best-first
    std::vector<AltHit> alts;
    {
      Coord c;
      size_t sub_idx = intervals.size();  // iterate .rev()
      CoordIter sub_it;
      bool sub_active = false;
      bool best_phase = true;
      while (alts.size() < 2) {
        bool got = false;
        if (best_phase) {
          if (best_it.next(&c)) got = true; else best_phase = false;
        }
        if (!got && !best_phase) {
          while (true) {
            if (sub_active) {
              if (sub_it.next(&c)) { got = true; break; }
              sub_active = false;
            }
            // advance to the next non-duplicate sub-optimal interval
            bool found = false;
            while (sub_idx > 0) {
              sub_idx -= 1;
              if (!interval_cross_check(best.interval, intervals[sub_idx].interval)) { found = true; break; }
            }
            if (!found) break;
            sub_it = interval2coordinate(intervals[sub_idx], ix, draw_u32(read_seed, draw_k++), ctr);
            if (!sub_it.pr.valid) continue;  // `.ok()` filter
            sub_active = true;
          }
        }
        if (!got) break;
        AltHit a;
        a.tid = c.tid; a.relative_pos = c.relative_pos; a.absolute_pos = c.absolute_pos; a.backward = c.backward;
        to_bam_fields(c.interval->edit_operations, c.backward, c.absolute_pos, ix, &a.cigar, &a.md, &a.nm);
        a.interval_size = c.interval->interval.size;
        a.score = c.interval->alignment_score;
        alts.push_back(a);
      }
    }
    for (const AltHit& a : alts) {
      rec.xa += ix.contigs[a.tid].name + "," + (a.backward ? "-" : "+") + std::to_string(a.relative_pos + 1) + "," +
                cigar_string(a.cigar) + "," + a.md + "," + std::to_string(a.nm) + "," + std::to_string(a.interval_size) +
                "," + format_score_2(a.score) + ";";
    }
    rec.alts = alts;
    rec.x0 = updated_size > (usize)INT32_MAX ? INT32_MAX : (int32_t)updated_size;
    {
      usize x1 = 0;
      for (const Hit& h : intervals)
        if (!interval_cross_check(best.interval, h.interval)) x1 += h.interval.size;
      rec.x1 = x1 > (usize)INT32_MAX ? INT32_MAX : (int32_t)x1;
    }
    rec.xs = intervals.empty() ? 0.0f : intervals.back().alignment_score;
    rec.xt = updated_size == 0 ? 'N' : (updated_size == 1 ? 'U' : 'R');
    rec.mapped = true;
    rec.tid = bc.tid; rec.pos = bc.relative_pos; rec.absolute_pos = bc.absolute_pos; rec.backward = bc.backward;
    rec.mapq = estimate_mapping_quality(best, updated_size, intervals, p);
    rec.alignment_score = best.alignment_score;
    rec.best_interval = best.interval;
    to_bam_fields(best.edit_operations, bc.backward, bc.absolute_pos, ix, &rec.cigar, &rec.md, &rec.nm);
    return rec;
  }
  rec.mapped = false;
  rec.mapq = 0;
  return rec;
}

}  // namespace ora
