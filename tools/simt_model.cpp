// tools/simt_model.cpp — ANALYSIS TOOL (not product, not a test).  Runs the product's per-read search logic
// (mapad_b200/csrc/search_core.cuh, compiled as plain C++ with MAPAD_STEP_STATS) for 32 reads "in lock step" the way
// k_search_pool's flat loop does — every lane expands one frame per iteration, a lane whose read finished takes the
// next one — and counts, in units of dependent memory round trips, what a warp pays for the data-dependent loops:
//   current    per iteration  1 + max_l(trickle_l) + 2 + sum_{i < max_l(n_cand_l)} (1 + max_l bubble_{l,i})
//   flat       per iteration  max_l [ 3 + trickle_l + sum_i (1 + bubble_{l,i}) ]      (micro-step state machine)
//   useful     sum over lanes of the per-lane chain
// so that useful / (32 x cost) is the lane efficiency each structure can reach at best.
#include <cstdio>
#include <cstring>
#include <vector>

#define MAPAD_STEP_STATS 1
#include "../include/mapad_gpu.h"
#include "../mapad_b200/csrc/dev_index_build.hpp"
#include "../mapad_b200/csrc/epilogue_core.cuh"
#include "../mapad_b200/csrc/host_index.hpp"
#include "../mapad_b200/csrc/host_params.hpp"

namespace mapad { thread_local StepStats* g_step_stats = nullptr; }
using namespace mapad;

// physical slot (in 8-byte entries) of heap index i under the three candidate layouts
static inline uint64_t slot_linear0(uint32_t i) { return i; }
static inline uint64_t slot_linear1(uint32_t i) { return (uint64_t)i + 1; }
// "family" layout: levels 0,1 in line 0; for k >= 1 the two children (level 2k) and four grandchildren (level 2k+1) of
// every node of level 2k-1 share one 64-byte line (6 of 8 slots used) — what one pop_max trickle-down step reads
static inline uint64_t slot_family(uint32_t i) {
  const uint32_t y = i + 1;
  const int lvl = 31 - __builtin_clz(y);
  if (lvl < 2) return i;
  const int odd = lvl & 1;
  const uint32_t o1 = y >> (1 + odd);
  const int lo = lvl - 1 - odd;
  const uint32_t j = o1 ^ (1u << lo);
  const uint32_t a = 0xAAAAAAAAu & ((1u << (lo - 1)) - 1u);
  const uint64_t line = 1ull + a + j;
  const uint32_t sl = odd ? 2u + (y & 3u) : (y & 1u);
  return line * 8 + sl;
}
template <class F>
static void count_lines(const StepStats& st, F slot, uint32_t hot_below, double& sectors, double& lines) {
  // distinct 32-byte sectors / 64-byte lines among the heap entries one expansion touched, ignoring the top of the heap
  // (indices < hot_below are assumed to stay cached)
  uint64_t sec[512], lin[512];
  int ns = 0, nl = 0;
  for (int k = 0; k < st.n_heap_idx; ++k) {
    if (st.heap_idx[k] < hot_below) continue;
    const uint64_t b = slot(st.heap_idx[k]) * 8;
    const uint64_t s_ = b >> 5, l_ = b >> 6;
    bool f = false;
    for (int q = 0; q < ns; ++q) if (sec[q] == s_) { f = true; break; }
    if (!f) sec[ns++] = s_;
    f = false;
    for (int q = 0; q < nl; ++q) if (lin[q] == l_) { f = true; break; }
    if (!f) lin[nl++] = l_;
  }
  sectors += ns; lines += nl;
}

struct Lane {
  std::vector<HeapEnt> heap;
  std::vector<NodeT<false>> nodes;
  std::vector<HitTmp> hits;
  Workspace<false> ws;
  SearchState<false> st;
  SearchCounters ctr;
  SearchJob job;
  bool have = false;
};

extern "C" int simt_model(const mapad_index* index, const mapad_params* params, const mapad_reads* in, uint32_t cap, double* out8, double* heap6) {
  const HostIndex* hix = reinterpret_cast<const HostIndex*>(index);
  IndexMeta meta;
  std::vector<uint8_t> blob;
  int rc = build_device_blob(*hix, meta, blob, 0);
  if (rc) return rc;
  if (meta.wide) return MAPAD_EINVAL;
  DevIndex ix{meta, blob.data()};
  BatchPrep bp;
  rc = prepare_batch(*params, *in, bp);
  if (rc) return rc;
  const DevParams& P = bp.dp;
  ReadBatch rb;
  rb.n_reads = in->n_reads; rb.seq = in->seq; rb.qual = in->qual; rb.offsets = in->offsets; rb.seeds = in->seeds;
  rb.starts = bp.starts.empty() ? nullptr : bp.starts.data();
  rb.custom_pen = nullptr;
  const uint64_t tb = bp.total_bases;
  std::vector<PenRow> delta(tb + 1);
  std::vector<float> dpen(tb + 1), dcomp(tb + 1);
  // prologue for all reads
  for (uint64_t r = 0; r < in->n_reads; ++r) {
    const uint64_t o = in->offsets[r];
    const int L = (int)(in->offsets[r + 1] - o);
    if (L == 0) continue;
    for (int j = 0; j < L; ++j) penalty_row(P, bp.qual_table, rb, o, j, L, delta.data(), dpen.data());
    const int split = alignment_start(P, rb, r, L);
    uint32_t dsteps = 0;
    for (int half = 0; half < 2; ++half) {
      const int part_len = half == 0 ? split : L - split;
      float* dout = dcomp.data() + o + (half == 0 ? 0 : split);
      if (part_len > 0) dout[0] = 0.0f;
      DScan sc[15];
      for (int l = 0; l < 15; ++l) dscan_init<false>(ix, sc[l], l);
      for (int idx = 0; idx + 1 < part_len; ++idx) {
        float v = 0.0f;
        for (int l = 0; l < 15; ++l) if (l <= idx) { dscan_step<false>(ix, sc[l], half, idx, L, rb.seq + o, dpen.data() + o, dsteps); v = fmin_rs(v, sc[l].z); }
        dout[idx + 1] = v;
      }
    }
  }
  std::vector<Lane> lanes(32);
  for (Lane& l : lanes) {
    l.heap.resize(cap); l.nodes.resize(cap); l.hits.resize(MAPAD_MAX_HITS);
    l.ws = Workspace<false>{l.heap.data(), l.nodes.data(), l.hits.data(), cap};
  }
  uint64_t next = 0;
  double iters = 0, cost_cur = 0, cost_flat = 0, useful = 0, frames = 0, cost_cur_push = 0, cost_cur_trickle = 0, skipped = 0;
  StepStats ss[32];
  while (true) {
    bool any = false;
    for (int l = 0; l < 32; ++l) {
      Lane& ln = lanes[l];
      ss[l].trickle = ss[l].pushes = ss[l].n_heap_idx = 0; memset(ss[l].bubble, 0, sizeof ss[l].bubble);
      ss[l].n_cand = -1;  // idle lane
      while (!ln.have && next < in->n_reads) {
        const uint64_t r = next++;
        const uint64_t o = in->offsets[r];
        const int L = (int)(in->offsets[r + 1] - o);
        if (L <= 0) continue;
        const int split = alignment_start(P, rb, r, L);
        ln.job = make_job(P, bp.bound_table.data(), rb.seq + o, L, split, delta.data() + o, dcomp.data() + o);
        if (search_begin<false>(ix, ln.job, ln.ws, ln.st, ln.ctr) == STEP_OVERFLOW) continue;
        ln.have = true;
      }
      if (!ln.have) continue;
      any = true;
      ss[l].n_cand = 0;
      g_step_stats = &ss[l];
      const int src = search_step<false>(ix, P, ln.job, ln.ws, ln.st, ln.ctr);
      g_step_stats = nullptr;
      frames += 1;
      if (heap6) {
        count_lines(ss[l], slot_linear0, 31, heap6[0], heap6[1]);
        count_lines(ss[l], slot_linear1, 31, heap6[2], heap6[3]);
        count_lines(ss[l], slot_family, 31, heap6[4], heap6[5]);
        heap6[6] += ss[l].n_heap_idx;
      }
      if (src != STEP_CONTINUE) { ln.have = false; if (src == STEP_OVERFLOW) skipped += 1; }
    }
    if (!any) break;
    iters += 1;
    int max_trickle = 0, max_cand = 0;
    int max_bubble[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double flat = 0;
    for (int l = 0; l < 32; ++l) {
      if (ss[l].n_cand < 0) continue;
      const int nc = ss[l].n_cand;
      if (ss[l].trickle > max_trickle) max_trickle = ss[l].trickle;
      if (nc > max_cand) max_cand = nc;
      double chain = 3 + ss[l].trickle;
      for (int i = 0; i < nc; ++i) {
        const int b = i < ss[l].pushes ? ss[l].bubble[i] : 0;
        if (b > max_bubble[i]) max_bubble[i] = b;
        chain += 1 + b;
      }
      useful += chain;
      if (chain > flat) flat = chain;
    }
    double cur = 3 + max_trickle;
    cost_cur_trickle += max_trickle;
    for (int i = 0; i < max_cand; ++i) { cur += 1 + max_bubble[i]; cost_cur_push += 1 + max_bubble[i]; }
    cost_cur += cur;
    cost_flat += flat;
  }
  out8[0] = iters; out8[1] = frames; out8[2] = cost_cur; out8[3] = cost_flat; out8[4] = useful; out8[5] = cost_cur_trickle; out8[6] = cost_cur_push;
  out8[7] = skipped;
  return 0;
}
