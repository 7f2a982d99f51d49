// tests/emu/emu_harness.cpp — TEST INFRASTRUCTURE.  Compiles the product's per-read device logic
// (mapad_b200/csrc/{dev_index,search_core,epilogue_core}.cuh) as plain C++ and runs it one "thread"
// at a time on the CPU, so that the non-GPU test-suite can check it against the oracle without a
// device.  It is NOT a product path: the C ABI in libmapad_gpu.so never runs without CUDA.
#include <cstring>
#include <vector>

#include "../../include/mapad_gpu.h"
#include "../../mapad_b200/csrc/dev_index_build.hpp"
#include "../../mapad_b200/csrc/epilogue_core.cuh"
#include "../../mapad_b200/csrc/host_index.hpp"
#include "../../mapad_b200/csrc/host_params.hpp"

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
