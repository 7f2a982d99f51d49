// tests/emu/emu_harness.cpp — TEST INFRASTRUCTURE.  Compiles the product's per-read device logic
// (mapad_b200/csrc/{dev_index,search_core,epilogue_core}.cuh) as plain C++ and runs it one "thread"
// at a time on the CPU, so that the non-GPU test-suite can check it against the oracle without a
// device.  It is NOT a product path: the C ABI in libmapad_gpu.so never runs without CUDA.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../include/mapad_gpu.h"
#include "../../mapad_b200/csrc/dev_index_build.hpp"
#include "../../mapad_b200/csrc/epilogue_core.cuh"
#include "../../mapad_b200/csrc/host_index.hpp"
#include "../../mapad_b200/csrc/host_params.hpp"
#include "../../mapad_b200/csrc/search_group.cuh"
#include "simt_emu.hpp"

using namespace mapad;

namespace {
struct EmuOut {
  std::vector<mapad_record> records;
  std::vector<mapad_hit> hits;
  std::vector<mapad_edit_op> ops;
  std::vector<uint32_t> cigar;
  std::vector<char> text;
};

template <bool WIDE>
int run(const DevIndex& ix, const BatchPrep& bp, const mapad_reads& in, uint32_t cap, EmuOut& out) {
  const DevParams& P = bp.dp;
  ReadBatch rb;
  rb.n_reads = in.n_reads; rb.seq = in.seq; rb.qual = in.qual; rb.offsets = in.offsets; rb.seeds = in.seeds;
  rb.starts = bp.starts.empty() ? nullptr : bp.starts.data();
  rb.custom_pen = in.custom_penalties ? in.custom_penalties : (bp.custom_pen.empty() ? nullptr : bp.custom_pen.data());
  const uint64_t tb = bp.total_bases;
  std::vector<PenRow> delta(tb + 1);
  std::vector<float> dpen(tb + 1), dcomp(tb + 1);
  std::vector<HeapEnt> heap(cap);
  std::vector<NodeT<WIDE>> nodes(cap);
  std::vector<HitTmp> hit_tmp(MAPAD_MAX_HITS);
  Workspace<WIDE> ws{heap.data(), nodes.data(), hit_tmp.data(), cap};
  out.records.assign(in.n_reads, mapad_record());
  out.cigar.assign(64 + 8 * in.n_reads + 4 * tb, 0);
  out.text.assign(64 + 16 * in.n_reads + 8 * tb, 0);
  uint32_t cig_cur = 0, text_cur = 0, overflow = 0;
  OutPools pools{out.cigar.data(), (uint32_t)out.cigar.size(), &cig_cur, out.text.data(), (uint32_t)out.text.size(), &text_cur, &overflow};
  const uint64_t base0 = in.n_reads ? in.offsets[0] : 0;
  for (uint64_t r = 0; r < in.n_reads; ++r) {
    const uint64_t o = in.offsets[r] - base0;
    const int L = (int)(in.offsets[r + 1] - in.offsets[r]);
    mapad_record& rec = out.records[r];
    memset(&rec, 0, sizeof rec);
    rec.tid = -1; rec.pos = -1;
    rec.hit_off = (uint32_t)out.hits.size();
    if (L == 0) continue;
    ReadBatch rb0 = rb;  // kernels index seq/qual by absolute offsets
    rb0.seq = in.seq + base0; rb0.qual = in.qual + base0;
    if (rb0.custom_pen && in.custom_penalties) rb0.custom_pen = in.custom_penalties + 4 * base0;
    for (int j = 0; j < L; ++j) penalty_row(P, bp.qual_table, rb0, o, j, L, delta.data(), dpen.data());
    const int split = alignment_start(P, rb, r, L);
    // D array: 15 lanes in lock step
    uint32_t dsteps = 0;
    for (int half = 0; half < 2; ++half) {
      const int part_len = half == 0 ? split : L - split;
      float* dout = dcomp.data() + o + (half == 0 ? 0 : split);
      if (part_len > 0) dout[0] = 0.0f;
      DScan sc[15];
      for (int l = 0; l < 15; ++l) dscan_init<WIDE>(ix, sc[l], l);
      for (int idx = 0; idx + 1 < part_len; ++idx) {
        float v = 0.0f;
        for (int l = 0; l < 15; ++l) {
          if (l <= idx) {
            dscan_step<WIDE>(ix, sc[l], half, idx, L, rb0.seq + o, dpen.data() + o, dsteps);
            v = fmin_rs(v, sc[l].z);
          }
        }
        dout[idx + 1] = v;
      }
    }
    SearchState<WIDE> st;
    SearchCounters ctr;
    int rc = search_read<WIDE>(ix, P, bp.bound_table.data(), rb0.seq + o, L, split, delta.data() + o, dcomp.data() + o, ws, st, ctr);
    if (rc != 0) return MAPAD_ELIMIT;
    rec.n_hits = st.n_hits;
    for (uint32_t h = 0; h < st.n_hits; ++h) {
      uint32_t n_left;
      uint32_t total = path_length<WIDE>(ws, ws.hits[h].node, split, n_left);
      mapad_hit mh;
      memset(&mh, 0, sizeof mh);
      mh.lower = ws.hits[h].lower; mh.lower_rev = ws.hits[h].lower_rev; mh.size = ws.hits[h].size;
      mh.alignment_score = ws.hits[h].score;
      mh.edit_off = (uint32_t)out.ops.size(); mh.edit_len = total;
      out.ops.resize(out.ops.size() + total);
      path_write<WIDE>(ws, ws.hits[h].node, split, total, n_left, out.ops.data() + mh.edit_off);
      out.hits.push_back(mh);
    }
    epilogue_read<WIDE>(ix, P, bp.bound_table.data(), L, in.seeds ? in.seeds[r] : 0u, out.hits.data() + rec.hit_off, rec.n_hits,
                        out.ops.data(), pools, rec);
    rec.n_hits = st.n_hits;
    rec.frames_popped = ctr.frames_popped;
    rec.d_ext_steps = dsteps;
    rec.flags = ctr.limit_hit ? 1u : 0u;
  }
  if (overflow) return MAPAD_ELIMIT;
  out.cigar.resize(cig_cur);
  out.text.resize(text_cur);
  return MAPAD_OK;
}

// ---- the group kernel (mapad_b200/csrc/search_group.cuh) under the SIMT emulator ----------------
struct GroupOpts {
  int group_size;       // G
  int n_groups;
  uint32_t pool_chunks; // total chunks incl. the 2 owned per group
  int lpt;              // 1: longest reads first
};

template <bool WIDE, int G>
void launch_groups(const GroupLaunch<WIDE>& a, int n_groups) {
  constexpr int TOPL = 11;
  std::vector<HeapEnt> smem((size_t)n_groups * TOPL * 8);
  if (G == 1) {
    for (int g = 0; g < n_groups; ++g) group_search_lane<WIDE, 1, TOPL>(a, (uint32_t)g, 0, smem.data() + (size_t)g * TOPL * 8);
    return;
  }
  simt_emu::run(n_groups, G, [&](int g, int l) { group_search_lane<WIDE, G, TOPL>(a, (uint32_t)g, l, smem.data() + (size_t)g * TOPL * 8); });
}

template <bool WIDE>
int run_group(const DevIndex& ix, const BatchPrep& bp, const mapad_reads& in, const GroupOpts& go, EmuOut& out, uint32_t* n_deferred_out) {
  const DevParams& P = bp.dp;
  const uint64_t n = in.n_reads;
  const uint64_t base0 = n ? in.offsets[0] : 0;
  std::vector<uint64_t> offs(n + 1);
  for (uint64_t r = 0; r <= n; ++r) offs[r] = n ? in.offsets[r] - base0 : 0;
  ReadBatch rb;
  rb.n_reads = n; rb.seq = in.seq + base0; rb.qual = in.qual + base0; rb.offsets = offs.data(); rb.seeds = in.seeds;
  rb.starts = bp.starts.empty() ? nullptr : bp.starts.data();
  rb.custom_pen = in.custom_penalties ? in.custom_penalties + 4 * base0 : (bp.custom_pen.empty() ? nullptr : bp.custom_pen.data());
  const uint64_t tb = bp.total_bases;
  std::vector<PenRow> delta(tb + 1);
  std::vector<float> dpen(tb + 1), dcomp(tb + 1);
  std::vector<uint32_t> dsteps(n, 0);
  for (uint64_t r = 0; r < n; ++r) {
    const uint64_t o = offs[r];
    const int L = (int)(offs[r + 1] - o);
    if (L == 0) continue;
    for (int j = 0; j < L; ++j) penalty_row(P, bp.qual_table, rb, o, j, L, delta.data(), dpen.data());
    const int split = alignment_start(P, rb, r, L);
    for (int half = 0; half < 2; ++half) {
      const int part_len = half == 0 ? split : L - split;
      float* dout = dcomp.data() + o + (half == 0 ? 0 : split);
      if (part_len > 0) dout[0] = 0.0f;
      DScan sc[15];
      for (int l = 0; l < 15; ++l) dscan_init<WIDE>(ix, sc[l], l);
      for (int idx = 0; idx + 1 < part_len; ++idx) {
        float v = 0.0f;
        for (int l = 0; l < 15; ++l)
          if (l <= idx) { dscan_step<WIDE>(ix, sc[l], half, idx, L, rb.seq + o, dpen.data() + o, dsteps[r]); v = fmin_rs(v, sc[l].z); }
        dout[idx + 1] = v;
      }
    }
  }
  // launch state
  const int ng = go.n_groups;
  const uint32_t n_chunks = go.pool_chunks;
  if (n_chunks < 4u) return MAPAD_EINVAL;
  std::vector<uint8_t> pool_mem((size_t)n_chunks * MAPAD_GCHUNK_BYTES + 64);
  std::vector<uint32_t> pool_next(n_chunks + 2);
  std::vector<unsigned long long> pool_heads((size_t)MAPAD_GPOOL_SHARDS * MAPAD_GPOOL_HEAD_STRIDE);
  gpool_init_host(n_chunks, pool_heads.data(), pool_next.data());  // every chunk is free; groups take their base chunks themselves
  GroupLaunch<WIDE> a;
  a.ix = ix; a.P = P; a.rb = rb; a.bound_table = bp.bound_table.data(); a.delta = delta.data(); a.dcomp = dcomp.data();
  a.pool.base = (uint8_t*)(((uintptr_t)pool_mem.data() + 63) & ~(uintptr_t)63);
  a.pool.n_chunks = n_chunks; a.pool.heads = pool_heads.data(); a.pool.next = pool_next.data();
  a.max_nodes = P.edit_tree_limit + 64; a.max_heap = P.stack_limit + 64;
  a.nt = (a.max_nodes >> (MAPAD_GCHUNK_SHIFT - 5)) + 1;
  a.ht = (heap_lines_for(a.max_heap) >> (MAPAD_GCHUNK_SHIFT - 6)) + 1;
  std::vector<uint32_t> tables((size_t)ng * (a.nt + a.ht));
  std::vector<HitTmp> hit_base((size_t)ng * MAPAD_MAX_HITS);
  a.tables = tables.data(); a.hit_base = hit_base.data();
  std::vector<uint32_t> work(n), deferred(n + 1);
  for (uint64_t r = 0; r < n; ++r) work[r] = (uint32_t)r;
  if (go.lpt) std::stable_sort(work.begin(), work.end(), [&](uint32_t x, uint32_t y) { return offs[x + 1] - offs[x] > offs[y + 1] - offs[y]; });
  a.work_list = work.data(); a.n_work = (uint32_t)n; a.deferred_list = deferred.data();
  Cursors cur;
  memset(&cur, 0, sizeof cur);
  a.cur = &cur;
  std::vector<ReadMid> mid(n);
  a.mid = mid.data();
  std::vector<mapad_hit> hit_pool(20 * n + 64);
  std::vector<mapad_edit_op> op_pool(64 + 20 * (tb + 8 * n));
  a.hit_pool = hit_pool.data(); a.hit_cap = (uint32_t)hit_pool.size();
  a.op_pool = op_pool.data(); a.op_cap = (uint32_t)op_pool.size();
  a.iter_budget = 0;
  a.flags_or = 0;
  a.prefetch = 3u;  // the prefetch address arithmetic runs (and is bounds-checked by the sanitizer builds); the prefetch itself is a no-op here
  // the host's retry loop (mapad_gpu.cu::search_with_groups): reads handed back because the pool ran dry are re-run
  // with fewer groups in flight
  uint32_t total_deferred = 0;
  int groups_now = ng;
  std::vector<uint32_t> work2;
  for (int attempt = 0;; ++attempt) {
    cur.queue_head = 0; cur.n_deferred = 0;
    gpool_init_host(n_chunks, pool_heads.data(), pool_next.data());
    a.flags_or = attempt ? 2u : 0u;
    a.patient = 0;
    switch (go.group_size) {
      case 1: launch_groups<WIDE, 1>(a, groups_now); break;
      case 2: launch_groups<WIDE, 2>(a, groups_now); break;
      case 4: launch_groups<WIDE, 4>(a, groups_now); break;
      case 8: launch_groups<WIDE, 8>(a, groups_now); break;
      case 16: launch_groups<WIDE, 16>(a, groups_now); break;
      case 32: launch_groups<WIDE, 32>(a, groups_now); break;
      default: return MAPAD_EINVAL;
    }
    if (cur.overflow & MAPAD_POOL_TIMEOUT_FLAG) return MAPAD_ELIMIT;
    if (cur.n_deferred == 0) break;
    total_deferred += cur.n_deferred;
    if (groups_now == 1 && cur.n_deferred >= a.n_work) { if (n_deferred_out) *n_deferred_out = total_deferred; return MAPAD_ELIMIT; }
    work2.assign(deferred.begin(), deferred.begin() + cur.n_deferred);
    a.work_list = work2.data(); a.n_work = cur.n_deferred;
    groups_now = groups_now > 1 ? groups_now / 2 : 1;
  }
  if (n_deferred_out) *n_deferred_out = total_deferred;
  if (cur.overflow) return MAPAD_ELIMIT;
  // epilogue
  out.records.assign(n, mapad_record());
  out.cigar.assign(64 + 8 * n + 4 * tb, 0);
  out.text.assign(64 + 16 * n + 8 * tb, 0);
  uint32_t cig_cur = 0, text_cur = 0, overflow = 0;
  OutPools pools{out.cigar.data(), (uint32_t)out.cigar.size(), &cig_cur, out.text.data(), (uint32_t)out.text.size(), &text_cur, &overflow};
  for (uint64_t r = 0; r < n; ++r) {
    const int L = (int)(offs[r + 1] - offs[r]);
    mapad_record& rec = out.records[r];
    memset(&rec, 0, sizeof rec);
    rec.tid = -1; rec.pos = -1;
    if (L == 0) continue;
    const ReadMid m = mid[r];
    epilogue_read<WIDE>(ix, P, bp.bound_table.data(), L, in.seeds ? in.seeds[r] : 0u, hit_pool.data() + m.hit_off, m.n_hits, op_pool.data(), pools, rec);
    rec.hit_off = m.hit_off; rec.n_hits = m.n_hits; rec.frames_popped = m.frames_popped; rec.d_ext_steps = dsteps[r]; rec.flags = m.flags;
  }
  if (overflow) return MAPAD_ELIMIT;
  out.hits.assign(hit_pool.begin(), hit_pool.begin() + cur.hit_cursor);
  out.ops.assign(op_pool.begin(), op_pool.begin() + cur.op_cursor);
  out.cigar.resize(cig_cur);
  out.text.resize(text_cur);
  return MAPAD_OK;
}
}  // namespace

extern "C" {
int emu_map_batch(const mapad_index* index, const mapad_params* params, const mapad_reads* in, uint32_t cap, int layout,
                  void** out_handle) {
  const HostIndex* hix = reinterpret_cast<const HostIndex*>(index);
  IndexMeta meta;
  std::vector<uint8_t> blob;
  int rc = build_device_blob(*hix, meta, blob, layout);
  if (rc) return rc;
  DevIndex ix{meta, blob.data()};
  BatchPrep bp;
  rc = prepare_batch(*params, *in, bp);
  if (rc) return rc;
  EmuOut* out = new EmuOut();
  rc = meta.wide ? run<true>(ix, bp, *in, cap, *out) : run<false>(ix, bp, *in, cap, *out);
  if (rc) { delete out; return rc; }
  *out_handle = out;
  return MAPAD_OK;
}
int emu_map_batch_group(const mapad_index* index, const mapad_params* params, const mapad_reads* in, int layout, int group_size,
                         int n_groups, uint32_t pool_chunks, int lpt, uint32_t* n_deferred_out, void** out_handle) {
  const HostIndex* hix = reinterpret_cast<const HostIndex*>(index);
  IndexMeta meta;
  std::vector<uint8_t> blob;
  int rc = build_device_blob(*hix, meta, blob, layout);
  if (rc) return rc;
  DevIndex ix{meta, blob.data()};
  BatchPrep bp;
  rc = prepare_batch(*params, *in, bp);
  if (rc) return rc;
  EmuOut* out = new EmuOut();
  GroupOpts go{group_size, n_groups, pool_chunks, lpt};
  rc = meta.wide ? run_group<true>(ix, bp, *in, go, *out, n_deferred_out) : run_group<false>(ix, bp, *in, go, *out, n_deferred_out);
  if (rc) { delete out; return rc; }
  *out_handle = out;
  return MAPAD_OK;
}
void emu_batch_view(void* h, mapad_results* r) {
  EmuOut* o = (EmuOut*)h;
  memset(r, 0, sizeof *r);
  r->n_reads = o->records.size(); r->records = o->records.data();
  r->hits = o->hits.data(); r->n_hits = o->hits.size();
  r->edit_ops = o->ops.data(); r->n_edit_ops = o->ops.size();
  r->cigar = o->cigar.data(); r->n_cigar = o->cigar.size();
  r->text = o->text.data(); r->n_text = o->text.size();
}
void emu_batch_free(void* h) { delete (EmuOut*)h; }
// rank queries through the device layout, for index cross-checks
int emu_occ4(const mapad_index* index, int layout, uint64_t row, uint64_t* out4, uint32_t* bwt_rank) {
  const HostIndex* hix = reinterpret_cast<const HostIndex*>(index);
  static thread_local const HostIndex* cached = nullptr;
  static thread_local int cached_layout = -2;
  static thread_local IndexMeta meta;
  static thread_local std::vector<uint8_t> blob;
  if (cached != hix || cached_layout != layout) { int rc = build_device_blob(*hix, meta, blob, layout); if (rc) return rc; cached = hix; cached_layout = layout; }
  DevIndex ix{meta, blob.data()};
  if (meta.wide) { occ4<true>(ix, row, out4); *bwt_rank = bwt_at<true>(ix, row); }
  else { occ4<false>(ix, row, out4); *bwt_rank = bwt_at<false>(ix, row); }
  return 0;
}
int emu_sa_get(const mapad_index* index, int layout, uint64_t row, uint64_t* out) {
  const HostIndex* hix = reinterpret_cast<const HostIndex*>(index);
  static thread_local const HostIndex* cached = nullptr;
  static thread_local int cached_layout = -2;
  static thread_local IndexMeta meta;
  static thread_local std::vector<uint8_t> blob;
  if (cached != hix || cached_layout != layout) { int rc = build_device_blob(*hix, meta, blob, layout); if (rc) return rc; cached = hix; cached_layout = layout; }
  DevIndex ix{meta, blob.data()};
  uint32_t steps = 0;
  *out = meta.wide ? sa_get<true>(ix, row, steps) : sa_get<false>(ix, row, steps);
  return 0;
}
float emu_libm(int fn, int iarg, float x) {
  switch (fn) {
    case 0: return emu::log2f_glibc(x);
    case 1: return emu::exp2f_glibc(x);
    case 2: return emu::log10f_glibc(x);
    default: return emu::powi_rt(x, iarg);
  }
}
void emu_libm_array(int fn, int iarg, uint64_t n, const float* in, float* out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = emu_libm(fn, iarg, in[i]);
}
}
