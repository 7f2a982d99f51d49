// search_warp.cuh — K2: warp-cooperative best-first k-mismatch search (device only).
//
// One warp owns one read at a time and pulls the next read from a global queue when it finishes.
//   * the min-max heap of (score, node) pairs lives in shared memory (HS entries per warp) and
//     spills its deep levels to a per-warp arena in HBM; the hit BinaryHeap is in shared memory too
//   * the edit tree doubles as frame storage: one 32 B (48 B wide) node per accepted child in HBM
//   * per popped frame: all lanes fetch the node and the two occ blocks (broadcast loads), lanes 0..8
//     evaluate the nine children (insertion, 4 x deletion, 4 x match/mismatch) in parallel,
//     a ballot collects the statically accepted ones, lane 0 replays them in the reference order
//     against the exact sequential heaps, then the accepted lanes store their nodes in parallel
//   * at the end of a read every hit is traced back by its own lane
// The sequential semantics (pop order among equal scores, hit order, slab key reuse, limits) are those
// of search_core.cuh::search_read, which restates /root/reference/src/map/mapping.rs:932-1383.
#pragma once
#include "search_core.cuh"

namespace mapad {

struct SplitHeapStore {  // heap index i < hs: shared memory, otherwise the HBM spill arena
  HeapEnt* sm;
  HeapEnt* gl;
  uint32_t hs;
  __device__ __forceinline__ HeapEnt get(uint32_t i) const { return i < hs ? sm[i] : gl[i - hs]; }
  __device__ __forceinline__ void set(uint32_t i, HeapEnt e) const { if (i < hs) sm[i] = e; else gl[i - hs] = e; }
};

struct HitEnt { float score; uint32_t node; };

__device__ __forceinline__ void hit_push(HitEnt* d, uint32_t& n, HitEnt x) {  // std BinaryHeap::push
  uint32_t pos = n;
  while (pos > 0) {
    uint32_t parent = (pos - 1) >> 1;
    HitEnt pe = d[parent];
    if (x.score <= pe.score) break;
    d[pos] = pe;
    pos = parent;
  }
  d[pos] = x;
  n += 1;
}

// MinMaxHeap::push by the whole warp (same result as search_core.cuh::mm_push).  The positions a new element can
// visit depend only on its index i: its parent p, then the grandparent chain of i (element stays on its level) or of p
// (element swapped with the parent).  Lane 0 fetches the parent, lanes 1..15 the chain of i, lanes 16..30 the chain
// of p — one memory latency for the whole bubble-up instead of one per level — a ballot finds where the climb stops,
// and the lanes whose ancestors move down write them in parallel.  `n` and `e` must be warp-uniform.
__device__ __forceinline__ void mm_push_warp(const SplitHeapStore& d, uint32_t& n, HeapEnt e, int lane) {
  const unsigned FULL = 0xffffffffu;
  const uint32_t i = n;
  n += 1;
  if (i == 0) {
    if (lane == 0) d.set(0, e);
    __syncwarp();
    return;
  }
  const uint32_t p = (i - 1) >> 1;
  const bool min_level = mm_on_min_level(i);
  const bool in_a = lane >= 1 && lane <= 15;
  const bool in_b = lane >= 16 && lane <= 30;
  const uint32_t lvl = in_a ? (uint32_t)lane : (in_b ? (uint32_t)(lane - 15) : 0u);   // 1-based level in the chain
  const uint32_t base1 = (in_b ? p : i) + 1u;                                          // 1-based heap index of the chain's origin
  const uint32_t anc1 = lvl ? base1 >> (2u * lvl) : 0u;                                // 1-based index of the lvl-th grandparent
  const bool valid = lvl != 0 && anc1 >= 1u;
  HeapEnt v = e;
  if (lane == 0) v = d.get(p);
  else if (valid) v = d.get(anc1 - 1u);
  const float pscore = __shfl_sync(FULL, v.score, 0);
  const uint32_t pnode = __shfl_sync(FULL, v.node, 0);
  const bool moved = min_level ? (e.score > pscore) : (e.score < pscore);
  const bool climb_max = min_level == moved;
  const bool wins = valid && (climb_max ? (e.score > v.score) : (e.score < v.score));
  const unsigned ball = __ballot_sync(FULL, wins);
  const unsigned chain = moved ? ((ball >> 16) & 0x7fffu) : ((ball >> 1) & 0x7fffu);
  const uint32_t t = (uint32_t)__ffs((int)~chain) - 1u;                                // leading run of winning levels
  const uint32_t cur = moved ? p : i;
  if (lane == 0) {
    if (moved) d.set(i, HeapEnt{pscore, pnode});
    const uint32_t fin = t == 0 ? cur : ((cur + 1u) >> (2u * t)) - 1u;
    d.set(fin, e);
  } else if (valid && (moved ? in_b : in_a) && lvl <= t) {
    const uint32_t below = lvl == 1 ? cur : (base1 >> (2u * (lvl - 1u))) - 1u;          // this ancestor moves one chain level down
    d.set(below, v);
  }
  __syncwarp();
}

// One level of MinMaxHeap::trickle_down at position i with the six candidates (2 children, 4 grandchildren) already in
// registers; same decisions as search_core.cuh::mm_trickle_down.  Returns true when the descent continues from the
// chosen grandchild (then `b` = which of the four).  Warp-uniform; lane 0 writes.
template <bool MAX>
__device__ __forceinline__ bool mm_trickle_level(const SplitHeapStore& d, uint32_t n, uint32_t& i, HeapEnt& e, const HeapEnt (&x)[6], int lane,
                                                 uint32_t& b) {
  const uint32_t c1 = 2 * i + 1, g1 = 4 * i + 3;
  uint32_t best = MAPAD_NO_NODE;
  float bk = e.score;
  HeapEnt be = e;
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const uint32_t idx = c < 2 ? c1 + c : g1 + (c - 2);
    if (idx < n && (MAX ? (x[c].score > bk) : (x[c].score < bk))) { best = idx; bk = x[c].score; be = x[c]; }
  }
  if (best == MAPAD_NO_NODE) return false;
  const bool was_child = best <= c1 + 1;
  if (lane == 0) d.set(i, be);
  i = best;
  if (was_child) return false;
  b = best - g1;
  const uint32_t p = (i - 1) >> 1;
  const HeapEnt pe = b < 2u ? x[0] : x[1];
  if (MAX ? (pe.score > e.score) : (pe.score < e.score)) {
    if (lane == 0) d.set(p, e);
    e = pe;
  }
  return true;
}

// MinMaxHeap::trickle_down by the whole warp.  The descent is a chain of dependent loads, one per level, and the deep
// levels of a large heap live in HBM/L2: lanes 0..29 therefore fetch the whole four-level subtree below the current
// position at once (2 + 4 + 8 + 16 entries), which covers the candidates of this level AND of the next one whichever
// grandchild is chosen — two levels per memory latency instead of one.
template <bool MAX>
__device__ __forceinline__ void mm_trickle_down_warp(const SplitHeapStore& d, uint32_t n, uint32_t i, int lane) {
  const unsigned FULL = 0xffffffffu;
  HeapEnt e = d.get(i);
  const uint32_t depth = lane < 2 ? 1u : (lane < 6 ? 2u : (lane < 14 ? 3u : 4u));
  const uint32_t off = lane < 2 ? (uint32_t)lane : (lane < 6 ? (uint32_t)lane - 2u : (lane < 14 ? (uint32_t)lane - 6u : (uint32_t)lane - 14u));
  while (true) {
    if (2 * i + 1 >= n) break;
    const uint32_t idx = ((i + 1u) << depth) - 1u + off;
    HeapEnt v = e;
    if (lane < 30 && idx < n) v = d.get(idx);
    HeapEnt x[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) { x[c].score = __shfl_sync(FULL, v.score, c); x[c].node = __shfl_sync(FULL, v.node, c); }
    uint32_t b = 0;
    if (!mm_trickle_level<MAX>(d, n, i, e, x, lane, b)) break;
    if (2 * i + 1 >= n) break;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const int src = c < 2 ? 6 + 2 * (int)b + c : 14 + 4 * (int)b + (c - 2);
      x[c].score = __shfl_sync(FULL, v.score, src);
      x[c].node = __shfl_sync(FULL, v.node, src);
    }
    if (!mm_trickle_level<MAX>(d, n, i, e, x, lane, b)) break;
    __syncwarp();  // lane 0's writes of these two levels precede the next fetch
  }
  if (lane == 0) d.set(i, e);
  __syncwarp();
}

template <bool WIDE>
__device__ __forceinline__ void node_store_w(NodeT<WIDE>* nodes, uint32_t id, const BiIv& iv, int start, int len, int gap_f, int gap_b,
                                             int ngaps, uint32_t parent, uint32_t op, uint32_t depth, uint32_t nleft) {
  NodeT<WIDE> n;
  n.parent = parent; n.op = op;
  n.lower = (decltype(n.lower))iv.lower; n.lower_rev = (decltype(n.lower))iv.lower_rev; n.size = (decltype(n.lower))iv.size;
  n.start = (int16_t)start; n.len = (int16_t)len;
  n.gap_f = (uint8_t)gap_f; n.gap_b = (uint8_t)gap_b; n.ngaps = (uint8_t)ngaps; n.pad0 = 0;
  n.pad1 = (depth > 0xffffu ? 0xffffu : depth) | ((nleft > 0xffffu ? 0xffffu : nleft) << 16);
  nodes[id] = n;
}

#define MAPAD_WARP_SMEM_EXTRA (MAPAD_MAX_HITS * 8 + 64)

template <bool WIDE>
__global__ void __launch_bounds__(128, 4)
k_search_warp(DevIndex ix, DevParams P, ReadBatch rb, const float* __restrict__ bound_table, const PenRow* __restrict__ delta,
              const float* __restrict__ dcomp, HeapEnt* gheap_base, NodeT<WIDE>* node_base, uint32_t cap, uint32_t hs,
              const uint32_t* __restrict__ work_list, uint32_t n_work, uint32_t* deferred_list, Cursors* cur, ReadMid* mid,
              mapad_hit* hit_pool, uint32_t hit_cap, mapad_edit_op* op_pool, uint32_t op_cap) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const uint64_t wslot = (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp_in_block;
  uint8_t* wbase = smem_raw + (size_t)warp_in_block * ((size_t)hs * sizeof(HeapEnt) + MAPAD_WARP_SMEM_EXTRA);
  HeapEnt* sm_heap = reinterpret_cast<HeapEnt*>(wbase);
  HitEnt* sm_hits = reinterpret_cast<HitEnt*>(wbase + (size_t)hs * sizeof(HeapEnt));
  uint32_t* sm_ids = reinterpret_cast<uint32_t*>(wbase + (size_t)hs * sizeof(HeapEnt) + MAPAD_MAX_HITS * 8);
  const SplitHeapStore heap{sm_heap, gheap_base + wslot * cap, hs};
  NodeT<WIDE>* nodes = node_base + wslot * cap;
  const float open_ext = fadd(P.gap_open, P.gap_extend);

  while (true) {
    uint32_t w = 0;
    if (lane == 0) w = atomicAdd(&cur->queue_head, 1u);
    w = __shfl_sync(FULL, w, 0);
    if (w >= n_work) break;
    const uint32_t r = work_list ? work_list[w] : w;
    const uint64_t o = rb.offsets[r];
    const int L = (int)(rb.offsets[r + 1] - o);
    ReadMid m;
    m.n_hits = 0; m.hit_off = 0; m.frames_popped = 0; m.flags = 0;
    if (L <= 0) {
      if (lane == 0) mid[r] = m;
      continue;
    }
    const int start_pos = alignment_start(P, rb, r, L);
    const uint8_t* seq = rb.seq + o;
    const PenRow* drow = delta + o;
    const float* dc = dcomp + o;
    const BoundCtx bc = bound_ctx(P, bound_table, L);
    // warp-uniform copies of the sequential state (lane 0 is authoritative)
    uint32_t heap_n = 0, node_hi = 1, free_head = MAPAD_NO_NODE, tree_len = 1, n_hits = 0;
    uint32_t frames = 0, limit_hit = 0;
    bool overflow = cap < 2;
    uint64_t best_size = 0;
    if (lane == 0 && !overflow) {
      node_store_w<WIDE>(nodes, 0, BiIv{0, 0, ix.m.n}, start_pos, 0, GAP_CLOSED, GAP_CLOSED, 0, 0, pack_op(0, MAPAD_ED_MATCH, 0), 0, 0);
      mm_push(heap, heap_n, HeapEnt{0.0f, 0});
    }
    __syncwarp();
    heap_n = __shfl_sync(FULL, heap_n, 0);

    while (!overflow) {
      // ---- select the maximum (all lanes; the heap top is always in shared memory, hs >= 3) ----
      if (heap_n == 0) break;
      const uint32_t mi = heap_n == 1 ? 0u : (heap_n == 2 ? 1u : (sm_heap[1].score > sm_heap[2].score ? 1u : 2u));
      const HeapEnt top = sm_heap[mi];
      frames += 1;
      // ---- frame: broadcast node load, issued before the heap is repaired so that both latencies overlap ----
      const NodeT<WIDE> pn = nodes[top.node];
      BiIv iv{(uint64_t)pn.lower, (uint64_t)pn.lower_rev, (uint64_t)pn.size};
      const int f_start = pn.start, f_len = pn.len, f_gap_f = pn.gap_f, f_gap_b = pn.gap_b, f_ngaps = pn.ngaps;
      const uint32_t p_depth = pn.pad1 & 0xffffu, p_nleft = pn.pad1 >> 16;
      const float score = top.score;
      int j, d_k, d_l;
      bool forward;
      if (f_start <= L - f_start - f_len) { j = f_start + f_len; forward = true; d_k = f_start; d_l = f_start + f_len; }
      else { j = f_start - 1; forward = false; d_k = f_start - 1; d_l = f_start + f_len - 1; }
      // ---- the two occ-block fetches of the four extensions go out now ... ----
      const BiIv ext_in = forward ? BiIv{iv.lower_rev, iv.lower, iv.size} : iv;
      ExtRaw<WIDE> ext_raw;
      extend_load<WIDE>(ix, ext_in, ext_raw);
      const PenRow row = drow[j];
      const float lower_bound = d_get(dc, L, start_pos, d_k, d_l);
      // ---- ... while lane 0 repairs the heap (MinMaxHeap::pop_max: move the last element into the hole, trickle down) ----
      __syncwarp();
      {
        const uint32_t n1 = heap_n - 1;
        if (mi < n1) {
          const HeapEnt last = heap.get(n1);
          if (lane == 0) heap.set(mi, last);
          __syncwarp();
          mm_trickle_down_warp<true>(heap, n1, mi, lane);
        }
      }
      heap_n -= 1;
      __syncwarp();
      const int side_gap = forward ? f_gap_f : f_gap_b;
      const float insertion_score = fadd(side_gap == GAP_INS ? P.gap_extend : open_ext, score);
      const float deletion_score = fadd(side_gap == GAP_DEL ? P.gap_extend : open_ext, score);
      const int gaps_open = side_gap == GAP_CLOSED ? f_ngaps + 1 : f_ngaps;
      if (n_hits > 0) {  // mapping.rs:1201-1208
        if (bound_reject_iterative(bc, fadd(score, lower_bound), sm_hits[0].score)) break;
      }
      BiIv ext[4];
      extend_finish<WIDE>(ix, ext_in, ext_raw, ext);
      // ---- lanes 0..8: one candidate child each (order: ins, del T, mm T, del G, mm G, del C, mm C, del A, mm A) ----
      const int c = lane;
      const int k = (c >= 1 && c <= 8) ? (c - 1) >> 1 : 0;
      const bool is_ins = c == 0;
      const bool is_del = c >= 1 && c <= 8 && ((c - 1) & 1) == 0;
      const bool is_mm = c >= 1 && c <= 8 && ((c - 1) & 1) == 1;
      BiIv ip = k == 0 ? ext[0] : (k == 1 ? ext[1] : (k == 2 ? ext[2] : ext[3]));
      const int rank = 4 - k;
      uint8_t cbase;
      int pen_idx;
      if (forward) { ip = BiIv{ip.lower_rev, ip.lower, ip.size}; cbase = complement_base(rank_base(rank)); pen_idx = 4 - rank; }
      else { cbase = rank_base(rank); pen_idx = rank - 1; }
      const float pen = pen_idx == 0 ? row.d[0] : (pen_idx == 1 ? row.d[1] : (pen_idx == 2 ? row.d[2] : row.d[3]));
      const float mm_score = fadd(pen, score);
      const int child_start = forward ? f_start : f_start - 1;
      const int dist_ins = j < L - j - 1 ? j : L - j - 1;
      const int dist5 = forward ? j : j + 1;
      const int dist3 = L - dist5;
      const int dist_del = dist5 < dist3 ? dist5 : dist3;
      bool accept = false;
      float my_score = 0.0f;
      BiIv my_iv = iv;
      int my_start = f_start, my_len = f_len, my_gf = f_gap_f, my_gb = f_gap_b, my_ng = f_ngaps;
      uint32_t my_op = 0;
      if (is_ins) {
        my_score = insertion_score;
        accept = !bound_reject(bc, fadd(insertion_score, lower_bound)) && dist_ins >= P.gap_dist_ends && gaps_open <= P.max_num_gaps_open;
        my_start = child_start; my_len = f_len + 1; my_ng = gaps_open;
        if (forward) my_gf = GAP_INS; else my_gb = GAP_INS;
        my_op = pack_op(j, MAPAD_ED_INSERTION, 0);
      } else if (is_del) {
        my_score = deletion_score;
        accept = ip.size >= 1 && !bound_reject(bc, fadd(deletion_score, lower_bound)) && dist_del >= P.gap_dist_ends &&
                 gaps_open <= P.max_num_gaps_open;
        my_iv = ip; my_ng = gaps_open;
        if (forward) my_gf = GAP_DEL; else my_gb = GAP_DEL;
        my_op = pack_op(j, MAPAD_ED_DELETION, cbase);
      } else if (is_mm) {
        my_score = mm_score;
        accept = ip.size >= 1 && !bound_reject(bc, fadd(mm_score, lower_bound)) && f_ngaps <= P.max_num_gaps_open;
        my_iv = ip; my_start = child_start; my_len = f_len + 1;
        if (forward) my_gf = GAP_CLOSED; else my_gb = GAP_CLOSED;
        my_op = cbase == seq[j] ? pack_op(j, MAPAD_ED_MATCH, 0) : pack_op(j, MAPAD_ED_MISMATCH, cbase);
      }
      unsigned mask = __ballot_sync(FULL, accept) & 0x1ffu;
      float sc[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) sc[q] = __shfl_sync(FULL, my_score, q);
      const bool grows = f_len + 1 == L;  // insertion / match children complete the read
      const uint32_t hits_before = n_hits;
      // ---- replay the accepted children in the reference order against the sequential heaps (check_and_push_stack_frame):
      //      every lane keeps the same copy of the sequential state, memory is written by lane 0, the heap push is cooperative ----
      if (mask) {
#pragma unroll
        for (int q = 0; q < 9; ++q) {
          if (!((mask >> q) & 1u)) continue;
          const float s = sc[q];
          if (n_hits > 0 && bound_reject_iterative(bc, s, sm_hits[0].score)) { mask &= ~(1u << q); continue; }
          uint32_t id;
          if (free_head != MAPAD_NO_NODE) { id = free_head; free_head = nodes[id].parent; }
          else {
            id = node_hi;
            if (id >= cap) { overflow = true; mask &= (1u << q) - 1u; break; }
            node_hi += 1;
          }
          tree_len += 1;
          if (lane == 0) sm_ids[q] = id;
          const bool is_deletion = q >= 1 && ((q - 1) & 1) == 0;
          if (grows && !is_deletion) {
            if (n_hits < MAPAD_MAX_HITS) {
              uint32_t nh = n_hits;
              if (lane == 0) hit_push(sm_hits, nh, HitEnt{s, id});
              n_hits += 1;
              __syncwarp();
            }
          } else {
            if (heap_n >= cap) { overflow = true; mask &= (1u << q) - 1u; break; }
            mm_push_warp(heap, heap_n, HeapEnt{s, id}, lane);
          }
        }
      }
      __syncwarp();
      // ---- accepted lanes store their nodes ----
      if (lane < 9 && ((mask >> lane) & 1u)) {
        const uint32_t left = (int)(my_op & 0xffffu) < start_pos ? 1u : 0u;
        node_store_w<WIDE>(nodes, sm_ids[lane], my_iv, my_start, my_len, my_gf, my_gb, my_ng, top.node, my_op, p_depth + 1, p_nleft + left);
      }
      __syncwarp();
      if (overflow) break;
      // ---- early exits (mapping.rs:1348-1355) ----
      if (n_hits > 9) break;
      if (n_hits != hits_before) best_size = (uint64_t)nodes[sm_hits[0].node].size;
      if (n_hits > 0 && best_size > 1) break;
      // ---- limits (mapping.rs:1358-1380) ----
      if (heap_n > P.stack_limit || tree_len > P.edit_tree_limit) {
        limit_hit += 1;
        if (P.stack_limit_abort) break;
        {  // MinMaxHeap::pop_min x excess, every lane keeping the same state (lane 0 writes, the descent is cooperative)
          long long e1 = (long long)heap_n - (long long)P.stack_limit;
          long long e2 = (long long)tree_len - (long long)P.edit_tree_limit;
          long long excess = e1 > e2 ? e1 : e2;
          for (long long e = 0; e < excess && heap_n > 0; ++e) {
            const HeapEnt lastv = heap.get(heap_n - 1);
            heap_n -= 1;
            HeapEnt mn = lastv;
            if (heap_n > 0) {
              mn = heap.get(0);
              __syncwarp();
              if (lane == 0) heap.set(0, lastv);
              __syncwarp();
              mm_trickle_down_warp<false>(heap, heap_n, 0, lane);
            }
            if (mn.node != 0) {  // Tree::remove (backtrack_tree.rs:49-53)
              if (lane == 0) nodes[mn.node].parent = free_head;
              free_head = mn.node;
              tree_len -= 1;
            }
          }
        }
        __syncwarp();
      }
    }
    if (overflow) {  // workspace too small: hand the read to the next lane
      if (lane == 0) deferred_list[atomicAdd(&cur->n_deferred, 1u)] = r;
      __syncwarp();
      continue;
    }
    // ---- emit hits: lane h traces hit h back to the root (extract_edit_operations, record.rs:465-500) ----
    m.frames_popped = frames;
    m.flags = (limit_hit ? 1u : 0u) | (work_list ? 2u : 0u);
    m.n_hits = n_hits;
    if (n_hits) {
      uint32_t my_node = 0, total = 0, n_left = 0;
      float my_hit_score = 0.0f;
      if ((uint32_t)lane < n_hits) {
        my_node = sm_hits[lane].node;
        my_hit_score = sm_hits[lane].score;
        const uint32_t pad = nodes[my_node].pad1;
        total = pad & 0xffffu; n_left = pad >> 16;
        if (total == 0xffffu || n_left == 0xffffu) total = path_length<WIDE>(PlainNodes<WIDE>{nodes}, my_node, start_pos, n_left);
      }
      uint32_t incl = total;  // inclusive warp scan of `total`
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t v = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += v;
      }
      const uint32_t sum = __shfl_sync(FULL, incl, 31);
      uint32_t hit_off = 0, op_base = 0;
      if (lane == 0) { hit_off = atomicAdd(&cur->hit_cursor, n_hits); op_base = atomicAdd(&cur->op_cursor, sum); }
      hit_off = __shfl_sync(FULL, hit_off, 0);
      op_base = __shfl_sync(FULL, op_base, 0);
      m.hit_off = hit_off;
      if ((uint32_t)lane < n_hits) {
        const uint32_t op_off = op_base + (incl - total);
        if ((uint64_t)op_off + total <= op_cap) path_write<WIDE>(PlainNodes<WIDE>{nodes}, my_node, start_pos, total, n_left, op_pool + op_off);
        else atomicOr(&cur->overflow, 1u);
        if ((uint64_t)hit_off + lane < hit_cap) {
          const NodeT<WIDE> hn = nodes[my_node];
          mapad_hit mh;
          mh.lower = hn.lower; mh.lower_rev = hn.lower_rev; mh.size = hn.size;
          mh.alignment_score = my_hit_score; mh.edit_off = op_off; mh.edit_len = total; mh.reserved = 0;
          hit_pool[hit_off + lane] = mh;
        } else {
          atomicOr(&cur->overflow, 1u);
        }
      }
    }
    if (lane == 0) mid[r] = m;
    __syncwarp();
  }
}

}  // namespace mapad
