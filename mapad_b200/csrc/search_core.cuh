// search_core.cuh — per-read logic of the hot path: penalty rows, D array, best-first search,
// hit extraction.  Dual-compilable (device / CPU emulation harness).
//
// Reference functions restated (all under /root/reference/src/map/):
//   compute_optimal_scores            mapping.rs:572-588
//   SimpleAncientDnaModel::get        sequence_difference_models.rs:117-207
//   get_min_penalty                   sequence_difference_models.rs:34-57
//   BiDArray::new / compute_part / get  bi_d_array.rs:24-224
//   k_mismatch_search                 mapping.rs:1012-1383
//   check_and_push_stack_frame        mapping.rs:932-987
//   MismatchBound::{reject, reject_iterative}  mismatch_bounds.rs:84-91,130-138,269-276
//   MinMaxHeap (min-max-heap crate), BinaryHeap (std), Tree/slab   
//   extract_edit_operations           record.rs:465-500
#pragma once
#include <cstdint>

#include "../../include/mapad_gpu.h"
#include "dev_index.cuh"
#include "libm_emu.cuh"

namespace mapad {

#define MAPAD_F32_LOWEST (-3.402823466e+38f)
#define MAPAD_F32_EPSILON (1.1920928955078125e-7f)

MAPAD_DEV float fmax_rs(float a, float b) { return a > b ? a : (b > a ? b : (a == a ? a : b)); }
MAPAD_DEV float fmin_rs(float a, float b) { return a < b ? a : (b < a ? b : (a == a ? a : b)); }
MAPAD_DEV float fma_rn(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return __builtin_fmaf(a, b, c);
#endif
}
MAPAD_DEV float fdiv_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b; return r;
#endif
}
using emu::fadd;
using emu::fmul;
MAPAD_DEV float fsub(float a, float b) { return fadd(a, -b); }

struct ReadBatch {  // device views of one batch
  uint64_t n_reads;
  const uint8_t* seq;
  const uint8_t* qual;
  const uint64_t* offsets;
  const uint32_t* seeds;
  const int16_t* starts;      // per-read alignment start (start_mode 2) or nullptr
  const float* custom_pen;    // MODEL_TABLE: get(i, len, b, read[i], q[i]) for b = A,C,G,T
};

MAPAD_DEV int alignment_start(const DevParams& P, const ReadBatch& rb, uint64_t read, int L) {
  if (P.start_mode == 0) return L;                      // SimpleAncientDnaModel (:209-211)
  if (P.start_mode == 1) return (int)((int16_t)L / 2);  // trait default (:59-61)
  return rb.starts[read];
}

// ---------------------------------------------------------------------------------------------
// Penalty rows.  For read position j: g[b] = sdm.get(j, L, b, read[j], q[j]) for b = A,C,G,T,
// opt = get_min_penalty(.., false), delta[b] = g[b] - opt, dpen = D-array penalty of position j.
// ---------------------------------------------------------------------------------------------
MAPAD_DEV void simple_model_row(const DevParams& P, const float* qual2prob, int j, int L, uint8_t to, uint8_t q, float g[4]) {
  const float seq_err = P.ignore_q ? P.default_q_prob : qual2prob[q];
  const float indep = fma_rn(seq_err, -P.divergence, fadd(seq_err, P.divergence));
  const float match_at = fma_rn(3.0f, -indep, 1.0f);
  const float l_indep = emu::log2f_glibc(fmax_rs(indep, MAPAD_F32_EPSILON));
  g[0] = g[1] = g[2] = g[3] = l_indep;
  if (to == 'A') {
    g[0] = emu::log2f_glibc(fmax_rs(match_at, MAPAD_F32_EPSILON));
  } else if (to == 'T') {
    g[3] = emu::log2f_glibc(fmax_rs(match_at, MAPAD_F32_EPSILON));
  }
  if (to == 'C' || to == 'T' || to == 'A' || to == 'G') {
    float p_fwd, p_rev;
    if (P.library == 0) {
      float a = emu::powi_rt(P.overhang5, j + 1);
      float b = emu::powi_rt(P.overhang3, (L - 1 - j) + 1);
      p_fwd = fma_rn(a, -b, fadd(a, b));
      p_rev = 0.0f;
    } else {
      p_fwd = emu::powi_rt(P.overhang5, j + 1);
      p_rev = emu::powi_rt(P.overhang5, (L - 1 - j) + 1);
    }
    if (to == 'C' || to == 'T') {
      float c_to_t = fma_rn(P.ss_rate, p_fwd, fmul(P.ds_rate, fsub(1.0f, p_fwd)));
      float v = to == 'C' ? fma_rn(fmul(4.0f, indep), c_to_t, fsub(match_at, c_to_t))
                          : fma_rn(fmul(4.0f, indep), -c_to_t, fadd(indep, c_to_t));
      g[1] = emu::log2f_glibc(fmax_rs(v, MAPAD_F32_EPSILON));
    } else {
      float g_to_a = fma_rn(P.ss_rate, p_rev, fmul(P.ds_rate, fsub(1.0f, p_rev)));
      float v = to == 'A' ? fma_rn(fmul(4.0f, indep), -g_to_a, fadd(indep, g_to_a))
                          : fma_rn(fmul(4.0f, indep), g_to_a, fsub(match_at, g_to_a));
      g[2] = emu::log2f_glibc(fmax_rs(v, MAPAD_F32_EPSILON));
    }
  }
}

struct PenRow { float d[4]; };  // delta for reference base A,C,G,T

MAPAD_DEV void penalty_row(const DevParams& P, const float* qual2prob, const ReadBatch& rb, uint64_t base_off, int j, int L,
                           PenRow* delta_out, float* dpen_out) {
  const uint8_t to = rb.seq[base_off + j];
  const uint8_t q = rb.qual[base_off + j];
  float g[4];
  if (P.model == MODEL_SIMPLE) simple_model_row(P, qual2prob, j, L, to, q, g);
  else {
    const float* cp = rb.custom_pen + 4 * (base_off + j);
    g[0] = cp[0]; g[1] = cp[1]; g[2] = cp[2]; g[3] = cp[3];
  }
  const int tr = base_rank(to);  // 0 if not ACGT
  float opt = 0.0f;
  if (tr != 0) {
    opt = MAPAD_F32_LOWEST;
    for (int b = 0; b < 4; ++b) opt = fmax_rs(opt, g[b]);
  }
  float best_mm = MAPAD_F32_LOWEST;
  for (int b = 0; b < 4; ++b)
    if (b + 1 != tr) best_mm = fmax_rs(best_mm, g[b]);
  PenRow row;
  for (int b = 0; b < 4; ++b) row.d[b] = fsub(g[b], opt);
  float mm_retval = fsub(best_mm, opt);
  int dist = j < L - 1 - j ? j : L - 1 - j;
  float dp = dist >= P.gap_dist_ends ? fmax_rs(mm_retval, P.gap_extend) : mm_retval;  // bi_d_array.rs:170-183
  delta_out[base_off + j] = row;
  dpen_out[base_off + j] = dp;
}

// ---------------------------------------------------------------------------------------------
// D array: one scan per offset 0..14 (bi_d_array.rs:104-198).  A "lane" owns one offset.
// ---------------------------------------------------------------------------------------------
struct DScan {
  BiIv iv;
  float z;
  int last_mm;
};
template <bool WIDE>
MAPAD_DEV void dscan_init(const DevIndex& ix, DScan& s, int offset) {
  s.iv = BiIv{0, 0, ix.m.n};
  s.z = 0.0f;
  s.last_mm = offset - 1;
}
// One step of scan `offset` at enumerate-index idx (idx >= offset).  half 0: backward D (scan the
// left part forward with forward_ext); half 1: forward D (scan the right part from the read end
// with backward_ext).  Positions are mapped to the full read like directed_index().
template <bool WIDE>
MAPAD_DEV void dscan_step(const DevIndex& ix, DScan& s, int half, int idx, int L, const uint8_t* seq, const float* dpen,
                          uint32_t& steps) {
  const int pos = half == 0 ? idx : L - 1 - idx;
  const int r = base_rank(seq[pos]);
  s.iv = half == 0 ? forward_ext_rank<WIDE>(ix, s.iv, r) : backward_ext_rank<WIDE>(ix, s.iv, r);
  steps += 1;
  if (s.iv.size < 1) {
    float m = MAPAD_F32_LOWEST;
    for (int j = s.last_mm + 1; j <= idx; ++j) m = fmax_rs(m, dpen[half == 0 ? j : L - 1 - j]);
    s.z = fadd(s.z, m);
    s.iv = BiIv{0, 0, ix.m.n};
    s.last_mm = idx;
  }
}

MAPAD_DEV float d_get(const float* dcomp, int L, int split, int backward_index, int forward_index) {  // bi_d_array.rs:200-224
  float d_rev = 0.0f, d_fwd = 0.0f;
  if (backward_index >= 0 && backward_index < L) d_rev = dcomp[backward_index];
  if (forward_index >= 0 && L >= 1 + forward_index) {
    int idx = L - (1 + forward_index) + split;
    if (idx < L) d_fwd = dcomp[idx];
  }
  return fadd(d_rev, d_fwd);
}

// ---------------------------------------------------------------------------------------------
// Search state containers
// ---------------------------------------------------------------------------------------------
struct alignas(8) HeapEnt { float score; uint32_t node; };  // 8-byte aligned: one 64-bit load / store per entry

enum { GAP_INS = 0, GAP_DEL = 1, GAP_CLOSED = 2 };

struct Frame {  // MismatchSearchStackFrame (mod.rs:105-115); the edit-tree node doubles as frame storage
  BiIv iv;
  int start, len;
  int gap_f, gap_b, ngaps;
  float score;
  uint32_t node;
};

template <bool WIDE> struct NodeT;
template <> struct alignas(16) NodeT<false> {  // 32 B = one sector
  uint32_t parent, op;
  uint32_t lower, lower_rev, size;
  int16_t start, len;
  uint8_t gap_f, gap_b, ngaps, pad0;
  uint32_t pad1;
};
template <> struct alignas(16) NodeT<true> {  // 48 B
  uint64_t lower, lower_rev, size;
  uint32_t parent, op;
  int16_t start, len;
  uint8_t gap_f, gap_b, ngaps, pad0;
  uint32_t pad1;
};

MAPAD_DEV uint32_t pack_op(int pos, int kind, uint8_t base) { return (uint32_t)pos | ((uint32_t)kind << 16) | ((uint32_t)base << 24); }

struct HitTmp {  // HitInterval before its edit operations are extracted
  float score;
  uint32_t node;
  uint64_t lower, lower_rev, size;
};

#define MAPAD_MAX_HITS 20   // the search returns once len() > 9, one expansion adds at most 9 (mapping.rs:1348)
#define MAPAD_NO_NODE 0xffffffffu

// Workspace policies.  A workspace provides: node(id) -> NodeT&, ensure_node(id) / ensure_heap(n) (grow or
// refuse), heap() -> a store with get/set, and the hit array.  `Workspace` is the contiguous arena of the sequential
// reference loop (emulation tests); GroupWorkspace (search_group.cuh) grows in chunks taken from the device-wide pool.
template <bool WIDE>
struct PlainNodes {
  NodeT<WIDE>* p;
  MAPAD_DEV NodeT<WIDE>& node(uint32_t id) const { return p[id]; }
};
template <bool WIDE>
struct Workspace {  // one per persistent thread; lives in global memory
  HeapEnt* heap_;
  NodeT<WIDE>* nodes;
  HitTmp* hits;
  uint32_t cap;   // capacity of heap_[] and nodes[]
  MAPAD_DEV NodeT<WIDE>& node(uint32_t id) const { return nodes[id]; }
  MAPAD_DEV bool ensure_node(uint32_t id) { return id < cap; }
  MAPAD_DEV bool ensure_heap(uint32_t n) { return n < cap; }
  MAPAD_DEV uint32_t min_cap() const { return cap; }
  struct Store {
    HeapEnt* d;
    MAPAD_DEV HeapEnt get(uint32_t i) const { return d[i]; }
    MAPAD_DEV void set(uint32_t i, HeapEnt e) const { d[i] = e; }
  };
  MAPAD_DEV Store heap() const { return Store{heap_}; }
};

struct SearchCounters { uint32_t frames_popped, tree_nodes, max_stack, limit_hit; };

// 32-byte node of the wide layout (text < 2^40 symbols): the three interval fields carry 40 bits each, so that one
// node is exactly one sector for both layouts (search_group.cuh).
struct alignas(16) NodeW32 {
  uint32_t parent, op;
  uint32_t lower_lo, lower_rev_lo, size_lo;
  uint8_t lower_hi, lower_rev_hi, size_hi, gaps;  // gaps = gap_f | gap_b << 2
  int16_t start, len;
  uint8_t ngaps, pad0;
  uint16_t pad1;
};

template <class N>
MAPAD_DEV void node_store(N& dst, const Frame& f, uint32_t parent, uint32_t op) {
  N n;
  n.parent = parent; n.op = op;
  n.lower = (decltype(n.lower))f.iv.lower; n.lower_rev = (decltype(n.lower))f.iv.lower_rev; n.size = (decltype(n.lower))f.iv.size;
  n.start = (int16_t)f.start; n.len = (int16_t)f.len;
  n.gap_f = (uint8_t)f.gap_f; n.gap_b = (uint8_t)f.gap_b; n.ngaps = (uint8_t)f.ngaps; n.pad0 = 0; n.pad1 = 0;
  dst = n;
}
MAPAD_DEV void node_store(NodeW32& dst, const Frame& f, uint32_t parent, uint32_t op) {
  NodeW32 n;
  n.parent = parent; n.op = op;
  n.lower_lo = (uint32_t)f.iv.lower; n.lower_rev_lo = (uint32_t)f.iv.lower_rev; n.size_lo = (uint32_t)f.iv.size;
  n.lower_hi = (uint8_t)(f.iv.lower >> 32); n.lower_rev_hi = (uint8_t)(f.iv.lower_rev >> 32); n.size_hi = (uint8_t)(f.iv.size >> 32);
  n.gaps = (uint8_t)(f.gap_f | (f.gap_b << 2));
  n.start = (int16_t)f.start; n.len = (int16_t)f.len;
  n.ngaps = (uint8_t)f.ngaps; n.pad0 = 0; n.pad1 = 0;
  dst = n;
}
// The 32 bytes of a node as they lie in memory (little endian), assembled in registers: building a node struct field by field
// made the compiler keep it in local memory (byte stores, 30 local stores and 266 L2 write sectors per popped frame in the
// ncu capture profiles/r2_g32_wide_50Mbp_raw.csv); a node is written with two 16-byte stores instead.
struct NodeWords { uint32_t w0, w1, w2, w3, w4, w5, w6, w7; };
MAPAD_DEV NodeWords node_words(const NodeT<false>*, const Frame& f, uint32_t parent, uint32_t op) {
  NodeWords v;
  v.w0 = parent; v.w1 = op;
  v.w2 = (uint32_t)f.iv.lower; v.w3 = (uint32_t)f.iv.lower_rev; v.w4 = (uint32_t)f.iv.size;
  v.w5 = ((uint32_t)f.start & 0xffffu) | ((uint32_t)f.len << 16);
  v.w6 = ((uint32_t)f.gap_f & 0xffu) | (((uint32_t)f.gap_b & 0xffu) << 8) | (((uint32_t)f.ngaps & 0xffu) << 16);
  v.w7 = 0u;
  return v;
}
MAPAD_DEV NodeWords node_words(const NodeW32*, const Frame& f, uint32_t parent, uint32_t op) {
  NodeWords v;
  v.w0 = parent; v.w1 = op;
  v.w2 = (uint32_t)f.iv.lower; v.w3 = (uint32_t)f.iv.lower_rev; v.w4 = (uint32_t)f.iv.size;
  v.w5 = ((uint32_t)(f.iv.lower >> 32) & 0xffu) | (((uint32_t)(f.iv.lower_rev >> 32) & 0xffu) << 8) |
         (((uint32_t)(f.iv.size >> 32) & 0xffu) << 16) | (((uint32_t)(f.gap_f | (f.gap_b << 2)) & 0xffu) << 24);
  v.w6 = ((uint32_t)f.start & 0xffffu) | ((uint32_t)f.len << 16);
  v.w7 = (uint32_t)f.ngaps & 0xffu;
  return v;
}
MAPAD_DEV void node_put(void* dst, const NodeWords& v) {  // dst is 16-byte aligned (nodes are alignas(16), 32 B)
#if defined(__CUDA_ARCH__)
  uint4* q = reinterpret_cast<uint4*>(dst);
  q[0] = make_uint4(v.w0, v.w1, v.w2, v.w3);
  q[1] = make_uint4(v.w4, v.w5, v.w6, v.w7);
#else
  uint32_t* q = reinterpret_cast<uint32_t*>(dst);
  q[0] = v.w0; q[1] = v.w1; q[2] = v.w2; q[3] = v.w3; q[4] = v.w4; q[5] = v.w5; q[6] = v.w6; q[7] = v.w7;
#endif
}

template <class N>
MAPAD_DEV void node_load(const N& src, uint32_t id, Frame& f) {
  N n = src;
  f.iv.lower = n.lower; f.iv.lower_rev = n.lower_rev; f.iv.size = n.size;
  f.start = n.start; f.len = n.len;
  f.gap_f = n.gap_f; f.gap_b = n.gap_b; f.ngaps = n.ngaps;
  f.node = id;
}
MAPAD_DEV void node_load(const NodeW32& src, uint32_t id, Frame& f) {
  NodeW32 n = src;
  f.iv.lower = (uint64_t)n.lower_lo | ((uint64_t)n.lower_hi << 32);
  f.iv.lower_rev = (uint64_t)n.lower_rev_lo | ((uint64_t)n.lower_rev_hi << 32);
  f.iv.size = (uint64_t)n.size_lo | ((uint64_t)n.size_hi << 32);
  f.start = n.start; f.len = n.len;
  f.gap_f = n.gaps & 3; f.gap_b = n.gaps >> 2; f.ngaps = n.ngaps;
  f.node = id;
}

// ---- min_max_heap::MinMaxHeap ------------------------------------------
// Generic over the backing store H (get(i) / set(i, e)): a plain array for the per-thread version,
// shared memory with a global-memory spill for the warp-cooperative kernel.
struct PlainHeapStore {
  HeapEnt* d;
  MAPAD_DEV HeapEnt get(uint32_t i) const { return d[i]; }
  MAPAD_DEV void set(uint32_t i, HeapEnt e) const { d[i] = e; }
};
MAPAD_DEV bool mm_on_min_level(uint32_t i) {
#if defined(__CUDA_ARCH__)
  int level = 31 - __clz((int)(i + 1));
#else
  int level = 31 - __builtin_clz(i + 1);
#endif
  return (level & 1) == 0;
}
template <class H>
MAPAD_DEV void mm_push(const H& d, uint32_t& n, HeapEnt e) {
  uint32_t i = n++;
  bool min_level = mm_on_min_level(i);
  bool climb_max;
  if (i > 0) {
    // The indices a new element can visit depend only on i: its parent p, then the grandparent chain of i or of p.
    // The parent and both first grandparents are fetched together, so the common case (stop after the parent and one
    // grandparent compare) costs one memory latency instead of two dependent ones.
    const uint32_t p = (i - 1) >> 1;
    const uint32_t gi = i >= 3 ? (p - 1) >> 1 : MAPAD_NO_NODE;                       // grandparent of i
    const uint32_t gp = p >= 3 ? (((p - 1) >> 1) - 1) >> 1 : MAPAD_NO_NODE;          // grandparent of p
    const HeapEnt pe = d.get(p);
    HeapEnt ge_i = e, ge_p = e;
    if (gi != MAPAD_NO_NODE) ge_i = d.get(gi);
    if (gp != MAPAD_NO_NODE) ge_p = d.get(gp);
    bool moved;
    if (min_level) {
      if (e.score > pe.score) { d.set(i, pe); i = p; climb_max = true; moved = true; } else { climb_max = false; moved = false; }
    } else {
      if (e.score < pe.score) { d.set(i, pe); i = p; climb_max = false; moved = true; } else { climb_max = true; moved = false; }
    }
    // first grandparent step from the prefetched values
    if (i >= 3) {
      const uint32_t g = moved ? gp : gi;
      const HeapEnt ge = moved ? ge_p : ge_i;
      if (climb_max ? (e.score > ge.score) : (e.score < ge.score)) { d.set(i, ge); i = g; }
      else { d.set(i, e); return; }
    }
  } else {
    climb_max = !min_level;
  }
  while (i >= 3) {
    uint32_t gp = (((i - 1) >> 1) - 1) >> 1;
    HeapEnt ge = d.get(gp);
    if (climb_max ? (e.score > ge.score) : (e.score < ge.score)) { d.set(i, ge); i = gp; } else break;
  }
  d.set(i, e);
}
template <bool MAX, class H>
MAPAD_DEV void mm_trickle_down(const H& d, uint32_t n, uint32_t i) {
  HeapEnt e = d.get(i);
  while (true) {
    const uint32_t c1 = 2 * i + 1;
    if (c1 >= n) break;
    const uint32_t g1 = 4 * i + 3;
    // the six candidates (2 children, 4 grandchildren) are fetched independently of each other so that a
    // level living in HBM costs one memory latency; candidate indices are increasing, so "stop at the first
    // index >= len" of the reference equals "skip the invalid ones"
    HeapEnt x[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const uint32_t idx = c < 2 ? c1 + c : g1 + (c - 2);
      x[c] = idx < n ? d.get(idx) : e;
    }
    uint32_t best = MAPAD_NO_NODE;
    float bk = e.score;
    HeapEnt be = e;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const uint32_t idx = c < 2 ? c1 + c : g1 + (c - 2);
      if (idx < n && (MAX ? (x[c].score > bk) : (x[c].score < bk))) { best = idx; bk = x[c].score; be = x[c]; }
    }
    if (best == MAPAD_NO_NODE) break;
    bool was_child = best <= c1 + 1;
    d.set(i, be);
    i = best;
    if (was_child) break;
    // the parent of the chosen grandchild is one of the two children fetched above (nothing wrote to it since)
    const uint32_t p = (i - 1) >> 1;
    const HeapEnt pe = (best - g1) < 2u ? x[0] : x[1];
    if (MAX ? (pe.score > e.score) : (pe.score < e.score)) { d.set(p, e); e = pe; }
  }
  d.set(i, e);
}
template <class H>
MAPAD_DEV bool mm_pop_max(const H& d, uint32_t& n, HeapEnt& out) {
  if (n == 0) return false;
  uint32_t m = n == 1 ? 0 : (n == 2 ? 1 : (d.get(1).score > d.get(2).score ? 1 : 2));
  HeapEnt item = d.get(n - 1);
  n -= 1;
  if (m < n) { HeapEnt t = d.get(m); d.set(m, item); item = t; mm_trickle_down<true>(d, n, m); }
  out = item;
  return true;
}
template <class H>
MAPAD_DEV bool mm_pop_min(const H& d, uint32_t& n, HeapEnt& out) {
  if (n == 0) return false;
  HeapEnt item = d.get(n - 1);
  n -= 1;
  if (n > 0) { HeapEnt t = d.get(0); d.set(0, item); item = t; mm_trickle_down<false>(d, n, 0); }
  out = item;
  return true;
}

// ---- std BinaryHeap<HitInterval> on HitTmp[] --------------------------
MAPAD_DEV void bh_sift_up(HitTmp* d, uint32_t start, uint32_t pos) {
  HitTmp e = d[pos];
  while (pos > start) {
    uint32_t parent = (pos - 1) >> 1;
    if (e.score <= d[parent].score) break;
    d[pos] = d[parent];
    pos = parent;
  }
  d[pos] = e;
}
MAPAD_DEV void bh_push(HitTmp* d, uint32_t& n, const HitTmp& x) { d[n] = x; bh_sift_up(d, 0, n); n += 1; }
// into_sorted_vec (ascending), in place
template <class T>
MAPAD_DEV void bh_into_sorted(T* d, uint32_t n) {
  uint32_t end = n;
  while (end > 1) {
    end -= 1;
    T t = d[0]; d[0] = d[end]; d[end] = t;
    uint32_t pos = 0;
    T e = d[0];
    uint32_t child = 1;
    bool done = false;
    while (child + 2 <= end) {
      if (d[child].score <= d[child + 1].score) child += 1;
      if (e.score >= d[child].score) { done = true; break; }
      d[pos] = d[child];
      pos = child;
      child = 2 * pos + 1;
    }
    if (!done && child + 1 == end && e.score < d[child].score) { d[pos] = d[child]; pos = child; }
    d[pos] = e;
  }
}

// ---- mismatch bound --------------------------------------------------------------------------
struct BoundCtx {
  int kind;
  float thr;      // discrete: k(L) * repr_mm; test: threshold
  float scale;    // continuous: L^exponent
  float cutoff;
  float repr_mm;
};
MAPAD_DEV BoundCtx bound_ctx(const DevParams& P, const float* bound_table, int L) {
  BoundCtx b;
  b.kind = P.bound_kind; b.repr_mm = P.repr_mm; b.cutoff = P.cutoff; b.scale = 1.0f; b.thr = 0.0f;
  float tv = (uint32_t)L < P.bound_table_len ? bound_table[L] : 0.0f;
  if (P.bound_kind == BOUND_DISCRETE) b.thr = fmul(tv, P.repr_mm);
  else if (P.bound_kind == BOUND_CONTINUOUS) b.scale = tv;
  else b.thr = P.test_threshold;
  return b;
}
// out of line on purpose: inlined, the compiler evaluates the division speculatively on every call of bound_reject
MAPAD_DEV_NOINLINE bool bound_reject_continuous(float value, float scale, float cutoff) { return fdiv_rn(value, scale) < cutoff; }
MAPAD_DEV bool bound_reject(const BoundCtx& b, float value) {
  if (b.kind == BOUND_CONTINUOUS) return bound_reject_continuous(value, b.scale, b.cutoff);
  return value < b.thr;
}
MAPAD_DEV bool bound_reject_iterative(const BoundCtx& b, float value, float reference) {
  if (b.kind == BOUND_TEST) return false;
  return value < fadd(reference, b.repr_mm);
}

// ---------------------------------------------------------------------------------------------
// k_mismatch_search for one read.  Returns 0 when finished, 1 when the workspace was too small
// (the caller re-runs the read in a lane with a larger workspace; nothing is emitted in that case).
// ---------------------------------------------------------------------------------------------
template <bool WIDE>
struct SearchState {
  uint32_t heap_n;
  uint32_t node_hi;      // slab: entries.len()
  uint32_t free_head;    // slab: next vacant key, MAPAD_NO_NODE if none
  uint32_t tree_len;     // slab: len()
  uint32_t n_hits;
  bool overflow;
};

template <bool WIDE, class WS>
MAPAD_DEV void check_and_push(WS& ws, SearchState<WIDE>& st, Frame f, uint32_t parent_node, uint32_t op, int L,
                              const BoundCtx& bc, const DevParams& P) {
  if (st.n_hits > 0) {
    if (bound_reject_iterative(bc, f.score, ws.hits[0].score)) return;
  }
  if (f.ngaps > P.max_num_gaps_open) return;
  // edit_tree.add_node (slab insert)
  uint32_t id;
  if (st.free_head != MAPAD_NO_NODE) {
    id = st.free_head;
    st.free_head = ws.node(id).parent;  // vacant entries chain through `parent`
  } else {
    id = st.node_hi;
    if (!ws.ensure_node(id)) { st.overflow = true; return; }
    st.node_hi += 1;
  }
  st.tree_len += 1;
  node_store(ws.node(id), f, parent_node, op);
  if (f.len == L) {
    HitTmp h;
    h.score = f.score; h.node = id; h.lower = f.iv.lower; h.lower_rev = f.iv.lower_rev; h.size = f.iv.size;
    if (st.n_hits < MAPAD_MAX_HITS) bh_push(ws.hits, st.n_hits, h);
    return;
  }
  if (!ws.ensure_heap(st.heap_n)) { st.overflow = true; return; }
  mm_push(ws.heap(), st.heap_n, HeapEnt{f.score, id});
}

// One read's constants and one pop-and-expand step of the search.  The thread kernels run
//   begin; while (step == STEP_CONTINUE);
// as a FLAT loop in which a thread that finishes a read picks up its next one, so that the 32 reads of a warp
// stay in the same loop body instead of waiting for the slowest read of a round.
#ifndef MAPAD_COMPACT_CAND
#define MAPAD_COMPACT_CAND 0
#endif
enum { STEP_CONTINUE = 0, STEP_DONE = 1, STEP_OVERFLOW = 2 };
struct SearchJob {
  const uint8_t* seq;
  const PenRow* delta;
  const float* dcomp;
  int L, start_pos;
  BoundCtx bc;
  float open_ext;
};
MAPAD_DEV SearchJob make_job(const DevParams& P, const float* bound_table, const uint8_t* seq, int L, int start_pos, const PenRow* delta,
                             const float* dcomp) {
  SearchJob job;
  job.seq = seq; job.delta = delta; job.dcomp = dcomp; job.L = L; job.start_pos = start_pos;
  job.bc = bound_ctx(P, bound_table, L);
  job.open_ext = fadd(P.gap_open, P.gap_extend);
  return job;
}

template <bool WIDE, class WS>
MAPAD_DEV int search_begin(const DevIndex& ix, const SearchJob& job, WS& ws, SearchState<WIDE>& st, SearchCounters& ctr) {
  st.heap_n = 0; st.node_hi = 0; st.free_head = MAPAD_NO_NODE; st.tree_len = 0; st.n_hits = 0; st.overflow = false;
  ctr.frames_popped = 0; ctr.tree_nodes = 0; ctr.max_stack = 0; ctr.limit_hit = 0;
  if (ws.min_cap() < 2 || !ws.ensure_node(0) || !ws.ensure_heap(0)) return STEP_OVERFLOW;
  Frame root;
  root.iv = BiIv{0, 0, ix.m.n};
  root.start = job.start_pos; root.len = 0; root.gap_f = GAP_CLOSED; root.gap_b = GAP_CLOSED; root.ngaps = 0;
  root.score = 0.0f; root.node = 0;
  node_store(ws.node(0), root, 0, pack_op(0, MAPAD_ED_MATCH, 0));  // tree.clear(): root = NodeId(0)
  st.node_hi = 1; st.tree_len = 1;
  mm_push(ws.heap(), st.heap_n, HeapEnt{0.0f, 0});
  return STEP_CONTINUE;
}

template <bool WIDE, class WS>
MAPAD_DEV int search_step(const DevIndex& ix, const DevParams& P, const SearchJob& job, WS& ws, SearchState<WIDE>& st, SearchCounters& ctr) {
  const int L = job.L;
  const BoundCtx& bc = job.bc;
  const float open_ext = job.open_ext;
  HeapEnt top;
  if (!mm_pop_max(ws.heap(), st.heap_n, top)) { ctr.tree_nodes = st.tree_len; return STEP_DONE; }
  ctr.frames_popped += 1;
  Frame sf;
  node_load(ws.node(top.node), top.node, sf);
  sf.score = top.score;
  int j, d_k, d_l;
  bool forward;
  if (sf.start <= L - sf.start - sf.len) {  // mapping.rs:1077-1097
    j = sf.start + sf.len; forward = true; d_k = sf.start; d_l = sf.start + sf.len;
  } else {
    j = sf.start - 1; forward = false; d_k = sf.start - 1; d_l = sf.start + sf.len - 1;
  }
  const PenRow row = job.delta[j];
  const int side_gap = forward ? sf.gap_f : sf.gap_b;
  const float insertion_score = fadd(side_gap == GAP_INS ? P.gap_extend : open_ext, sf.score);
  const float deletion_score = fadd(side_gap == GAP_DEL ? P.gap_extend : open_ext, sf.score);
  const int num_gaps_open = side_gap == GAP_CLOSED ? sf.ngaps + 1 : sf.ngaps;
  const float lower_bound = d_get(job.dcomp, L, job.start_pos, d_k, d_l);
  if (st.n_hits > 0) {  // mapping.rs:1201-1208
    if (bound_reject_iterative(bc, fadd(sf.score, lower_bound), ws.hits[0].score)) return STEP_DONE;
  }
  const int child_start = forward ? sf.start : sf.start - 1;
  // The (at most nine) children that pass the static tests are first collected, in the reference's order
  // (insertion; then for T,G,C,A: deletion, match/mismatch), and then replayed through ONE copy of
  // check_and_push in a run-time loop: with 32 independent reads per warp this keeps the threads in the same
  // instructions instead of nine separately predicated copies of the push code.
#if MAPAD_COMPACT_CAND
  // Variant (round-2 candidate, off by default): the nine candidates are kept as 4-bit codes (type, base index) in one
  // register and rebuilt from the popped frame in the push loop, instead of nine 56-byte frames in local memory — the
  // per-thread stack of k_search_pool (760 B x 512 threads per SM) is larger than L1 and competes with the heaps.
  uint64_t codes = 0;
  int n_cand = 0;
  {  // insertion (mapping.rs:1213-1242)
    int dist = j < L - j - 1 ? j : L - j - 1;
    if (!bound_reject(bc, fadd(insertion_score, lower_bound)) && dist >= P.gap_dist_ends) { codes |= 0ull << (4 * n_cand); n_cand += 1; }
  }
  BiIv ext[4];
  {
    BiIv in = forward ? BiIv{sf.iv.lower_rev, sf.iv.lower, sf.iv.size} : sf.iv;
    extend_all<WIDE>(ix, in, ext);
  }
  const bool del_ok = !bound_reject(bc, fadd(deletion_score, lower_bound));
  const int dist5 = forward ? j : j + 1;
  const int dist3 = L - dist5;
  const bool del_dist_ok = (dist5 < dist3 ? dist5 : dist3) >= P.gap_dist_ends;
  const uint8_t read_base = job.seq[j];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (ext[k].size < 1) continue;
    const int pen_idx = forward ? k : 3 - k;  // rank = 4 - k; forward: 4 - rank, backward: rank - 1
    if (del_ok && del_dist_ok) { codes |= (uint64_t)(4 | k) << (4 * n_cand); n_cand += 1; }
    const float mm_score = fadd(row.d[pen_idx], sf.score);
    if (!bound_reject(bc, fadd(mm_score, lower_bound))) { codes |= (uint64_t)(8 | k) << (4 * n_cand); n_cand += 1; }
  }
  for (int i = 0; i < n_cand && !st.overflow; ++i) {
    const uint32_t code = (uint32_t)(codes >> (4 * i)) & 15u;
    const uint32_t type = code >> 2, k = code & 3u;
    Frame ch = sf;
    uint32_t op;
    if (type == 0) {
      ch.start = child_start; ch.len = sf.len + 1;
      if (forward) ch.gap_f = GAP_INS; else ch.gap_b = GAP_INS;
      ch.score = insertion_score; ch.ngaps = num_gaps_open;
      op = pack_op(j, MAPAD_ED_INSERTION, 0);
    } else {
      BiIv ip = k == 0 ? ext[0] : (k == 1 ? ext[1] : (k == 2 ? ext[2] : ext[3]));
      const int rank = 4 - (int)k;
      uint8_t c;
      if (forward) { ip = BiIv{ip.lower_rev, ip.lower, ip.size}; c = complement_base(rank_base(rank)); }
      else c = rank_base(rank);
      ch.iv = ip;
      if (type == 1) {
        if (forward) ch.gap_f = GAP_DEL; else ch.gap_b = GAP_DEL;
        ch.score = deletion_score; ch.ngaps = num_gaps_open;
        op = pack_op(j, MAPAD_ED_DELETION, c);
      } else {
        const int pen_idx = forward ? (int)k : 3 - (int)k;
        const float pen = pen_idx == 0 ? row.d[0] : (pen_idx == 1 ? row.d[1] : (pen_idx == 2 ? row.d[2] : row.d[3]));
        ch.start = child_start; ch.len = sf.len + 1;
        if (forward) ch.gap_f = GAP_CLOSED; else ch.gap_b = GAP_CLOSED;
        ch.score = fadd(pen, sf.score);
        op = c == read_base ? pack_op(j, MAPAD_ED_MATCH, 0) : pack_op(j, MAPAD_ED_MISMATCH, c);
      }
    }
    check_and_push<WIDE, WS>(ws, st, ch, sf.node, op, L, bc, P);
  }
#else
  Frame cand[9];
  uint32_t cand_op[9];
  int n_cand = 0;
  // insertion (mapping.rs:1213-1242)
  {
    int dist = j < L - j - 1 ? j : L - j - 1;
    if (!bound_reject(bc, fadd(insertion_score, lower_bound)) && dist >= P.gap_dist_ends) {
      Frame c = sf;
      c.start = child_start; c.len = sf.len + 1;
      if (forward) c.gap_f = GAP_INS; else c.gap_b = GAP_INS;
      c.score = insertion_score; c.ngaps = num_gaps_open;
      cand[n_cand] = c; cand_op[n_cand] = pack_op(j, MAPAD_ED_INSERTION, 0); n_cand += 1;
    }
  }
  // bidirectional extension (mapping.rs:1245-1339)
  BiIv ext[4];
  {
    BiIv in = forward ? BiIv{sf.iv.lower_rev, sf.iv.lower, sf.iv.size} : sf.iv;
    extend_all<WIDE>(ix, in, ext);
  }
  const bool del_ok = !bound_reject(bc, fadd(deletion_score, lower_bound));
  const int dist5 = forward ? j : j + 1;
  const int dist3 = L - dist5;
  const bool del_dist_ok = (dist5 < dist3 ? dist5 : dist3) >= P.gap_dist_ends;
  const uint8_t read_base = job.seq[j];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    BiIv ip = ext[k];
    if (ip.size < 1) continue;
    const int rank = 4 - k;
    uint8_t c;
    int pen_idx;  // index of the reference base in PenRow
    if (forward) {
      ip = BiIv{ip.lower_rev, ip.lower, ip.size};
      c = complement_base(rank_base(rank));
      pen_idx = 4 - rank;  // complement: rank r -> 5 - r, index = 4 - r
    } else {
      c = rank_base(rank);
      pen_idx = rank - 1;
    }
    if (del_ok && del_dist_ok) {  // deletion (mapping.rs:1265-1302)
      Frame ch = sf;
      ch.iv = ip;
      if (forward) ch.gap_f = GAP_DEL; else ch.gap_b = GAP_DEL;
      ch.score = deletion_score; ch.ngaps = num_gaps_open;
      cand[n_cand] = ch; cand_op[n_cand] = pack_op(j, MAPAD_ED_DELETION, c); n_cand += 1;
    }
    const float mm_score = fadd(row.d[pen_idx], sf.score);
    if (!bound_reject(bc, fadd(mm_score, lower_bound))) {  // match / mismatch (mapping.rs:1307-1338)
      Frame ch = sf;
      ch.iv = ip;
      ch.start = child_start; ch.len = sf.len + 1;
      if (forward) ch.gap_f = GAP_CLOSED; else ch.gap_b = GAP_CLOSED;
      ch.score = mm_score;
      cand[n_cand] = ch;
      cand_op[n_cand] = c == read_base ? pack_op(j, MAPAD_ED_MATCH, 0) : pack_op(j, MAPAD_ED_MISMATCH, c);
      n_cand += 1;
    }
  }
  for (int i = 0; i < n_cand && !st.overflow; ++i) check_and_push<WIDE, WS>(ws, st, cand[i], sf.node, cand_op[i], L, bc, P);
#endif
  if (st.overflow) return STEP_OVERFLOW;
  if (st.heap_n > ctr.max_stack) ctr.max_stack = st.heap_n;
  // early exits (mapping.rs:1348-1355)
  if (st.n_hits > 9 || (st.n_hits > 0 && ws.hits[0].size > 1)) return STEP_DONE;
  // limits (mapping.rs:1358-1380)
  if (st.heap_n > P.stack_limit || st.tree_len > P.edit_tree_limit) {
    ctr.limit_hit += 1;
    if (P.stack_limit_abort) return STEP_DONE;
    long long e1 = (long long)st.heap_n - (long long)P.stack_limit;
    long long e2 = (long long)st.tree_len - (long long)P.edit_tree_limit;
    long long excess = e1 > e2 ? e1 : e2;
    for (long long e = 0; e < excess; ++e) {
      HeapEnt mn;
      if (mm_pop_min(ws.heap(), st.heap_n, mn)) {
        if (mn.node != 0) {  // Tree::remove (backtrack_tree.rs:49-53)
          ws.node(mn.node).parent = st.free_head;
          st.free_head = mn.node;
          st.tree_len -= 1;
        }
      }
    }
  }
  return STEP_CONTINUE;
}

template <bool WIDE, class WS>
MAPAD_DEV int search_read(const DevIndex& ix, const DevParams& P, const float* bound_table, const uint8_t* seq, int L, int start_pos,
                          const PenRow* delta, const float* dcomp, WS& ws, SearchState<WIDE>& st,
                          SearchCounters& ctr) {
  const SearchJob job = make_job(P, bound_table, seq, L, start_pos, delta, dcomp);
  int rc = search_begin<WIDE, WS>(ix, job, ws, st, ctr);
  while (rc == STEP_CONTINUE) rc = search_step<WIDE, WS>(ix, P, job, ws, st, ctr);
  ctr.tree_nodes = st.tree_len;
  return rc == STEP_OVERFLOW ? 1 : 0;
}

// extract_edit_operations (record.rs:465-500).  For the search order used here the bucket order of
// the reference collapses to: operations left of the start point in leaf->root order, then the
// others in root->leaf order.  `out` must hold `total` entries; returns eff_len via reference.
template <bool WIDE, class A>
MAPAD_DEV uint32_t path_length(const A& nodes, uint32_t node, int start_pos, uint32_t& n_left) {
  uint32_t total = 0;
  n_left = 0;
  while (node != 0) {
    const auto& nd = nodes.node(node);
    uint32_t op = nd.op;
    total += 1;
    if ((int)(op & 0xffffu) < start_pos) n_left += 1;
    node = nd.parent;
  }
  return total;
}
template <bool WIDE, class A>
MAPAD_DEV void path_write(const A& nodes, uint32_t node, int start_pos, uint32_t total, uint32_t n_left, mapad_edit_op* out) {
  uint32_t li = 0, ri = total;
  (void)n_left;
  while (node != 0) {
    const auto& nd = nodes.node(node);
    uint32_t op = nd.op;
    mapad_edit_op e;
    e.pos = (uint16_t)(op & 0xffffu); e.kind = (uint8_t)((op >> 16) & 0xffu); e.base = (uint8_t)(op >> 24);
    if ((int)e.pos < start_pos) out[li++] = e; else out[--ri] = e;
    node = nd.parent;
  }
}

}  // namespace mapad
