// mapad_gpu.cu — CUDA kernels (sm_100a) and the C ABI of include/mapad_gpu.h.
//
// Kernel inventory (SURVEY.md §2 "New sm_100a kernels"):
//   k_penalties   K1a  penalty rows: SimpleAncientDnaModel::get on the device, optimal scores
//   k_darray      K1b  BiDArray::new: 15 offset scans per read, one scan per lane of a 16-lane group
//   k_search_group K2  k_mismatch_search (search_group.cuh): persistent lane groups (G lanes per read), one flat loop, dynamic
//                      read queue (longest reads first), min-max heap in 64-byte family lines (top in shared memory) and
//                      edit tree in 256 KiB chunks of a device-wide pool; reads are never restarted
//   k_epilogue    K3   intervals_to_bam: best hit, SA locate, strand / contig, MAPQ, CIGAR / MD / NM, alts
//   k_gather      roofline denominator: independent random sector gathers
// There is no CPU fallback: every entry point fails with MAPAD_ENODEV without a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/mapad_gpu.h"
#include "dev_index_build.hpp"
#include "epilogue_core.cuh"
#include "search_group.cuh"
#include "host_index.hpp"
#include "host_params.hpp"

using namespace mapad;

// =================================================================================================
// kernels
// =================================================================================================
__global__ void __launch_bounds__(256) k_penalties(DevParams P, ReadBatch rb, const float* __restrict__ qual2prob,
                                                   PenRow* __restrict__ delta, float* __restrict__ dpen) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint64_t r = warp; r < rb.n_reads; r += n_warps) {
    const uint64_t o = rb.offsets[r];
    const int L = (int)(rb.offsets[r + 1] - o);
    for (int j = lane; j < L; j += 32) penalty_row(P, qual2prob, rb, o, j, L, delta, dpen);
  }
}

template <bool WIDE>
__global__ void __launch_bounds__(256) k_darray(DevIndex ix, DevParams P, ReadBatch rb, const float* __restrict__ dpen,
                                                float* __restrict__ dcomp, uint32_t* __restrict__ d_steps) {
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t group = gid >> 4, l16 = gid & 15;
  const uint32_t n_groups = (gridDim.x * blockDim.x) >> 4;
  const unsigned mask = 0xffffu << (threadIdx.x & 16);
  for (uint64_t r = group; r < rb.n_reads; r += n_groups) {
    const uint64_t o = rb.offsets[r];
    const int L = (int)(rb.offsets[r + 1] - o);
    const int split = L > 0 ? alignment_start(P, rb, r, L) : 0;
    uint32_t steps = 0;
    for (int half = 0; half < 2; ++half) {
      const int part_len = half == 0 ? split : L - split;
      float* dout = dcomp + o + (half == 0 ? 0 : split);
      if (l16 == 0 && part_len > 0) dout[0] = 0.0f;
      DScan sc;
      dscan_init<WIDE>(ix, sc, (int)l16);
      for (int idx = 0; idx + 1 < part_len; ++idx) {
        float v = 0.0f;
        if (l16 < 15 && (int)l16 <= idx) {
          dscan_step<WIDE>(ix, sc, half, idx, L, rb.seq + o, dpen + o, steps);
          v = sc.z;
        }
#pragma unroll
        for (int d = 8; d >= 1; d >>= 1) v = fmin_rs(v, __shfl_xor_sync(mask, v, d, 16));
        if (l16 == 0) dout[idx + 1] = v;
      }
    }
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) steps += __shfl_xor_sync(mask, steps, d, 16);
    if (l16 == 0) d_steps[r] = steps;
  }
}

template <bool WIDE>
__global__ void __launch_bounds__(128) k_epilogue(DevIndex ix, DevParams P, ReadBatch rb, const float* __restrict__ bound_table,
                                                  const ReadMid* __restrict__ mid, const uint32_t* __restrict__ d_steps,
                                                  const mapad_hit* __restrict__ hit_pool, const mapad_edit_op* __restrict__ op_pool,
                                                  OutPools pools, mapad_record* __restrict__ records) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rb.n_reads) return;
  const ReadMid m = mid[r];
  const int L = (int)(rb.offsets[r + 1] - rb.offsets[r]);
  mapad_record rec;
  rec.lf_steps = 0;
  epilogue_read<WIDE>(ix, P, bound_table, L, rb.seeds ? rb.seeds[r] : 0u, hit_pool + m.hit_off, m.n_hits, op_pool, pools, rec);
  rec.hit_off = m.hit_off;
  rec.n_hits = m.n_hits;
  rec.frames_popped = m.frames_popped;
  rec.d_ext_steps = d_steps[r];
  rec.flags = m.flags;
  records[r] = rec;
}

__global__ void k_debug_libm(int fn, int iarg, uint64_t n, const float* __restrict__ in, float* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = in[i];
  float y;
  switch (fn) {
    case 0: y = emu::log2f_glibc(x); break;
    case 1: y = emu::exp2f_glibc(x); break;
    case 2: y = emu::log10f_glibc(x); break;
    default: y = emu::powi_rt(x, iarg); break;
  }
  out[i] = y;
}

// Roofline denominator: every thread issues independent, uniformly random, `V`*16-byte loads.
template <int V>
__global__ void __launch_bounds__(256) k_gather(const uint4* __restrict__ table, uint64_t n_units, uint64_t per_thread, uint64_t seed,
                                                unsigned long long* sink) {
  uint64_t x = seed + (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull;
  uint32_t acc = 0;
  for (uint64_t i = 0; i < per_thread; ++i) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;  // xorshift64
    const uint64_t u = __umul64hi(x, n_units);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      uint4 t = __ldg(table + u * V + v);
      acc += t.x ^ t.y ^ t.z ^ t.w;
    }
  }
  if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

// =================================================================================================
// host side
// =================================================================================================
namespace {

template <class T>
struct DevBuf {  // grow-only device buffer
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n, bool exact = false) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = exact ? n : n + n / 4 + 64;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
template <class T>
struct PinBuf {  // grow-only pinned host buffer
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = n + n / 4 + 64;
    cudaError_t e = cudaMallocHost(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace

// Device-wide search workspace: ONE pool of 256 KiB chunks per GPU, shared by every handle (and every launch in flight) of
// the process — the per-read heaps and edit trees of all chunks in flight grow in it, so that memory goes where the heavy
// reads are instead of being partitioned per handle.  Created with the first handle on a device, freed with the last.
namespace {
struct DeviceArena {
  uint8_t* base = nullptr;
  uint32_t* next = nullptr;   // next pointers of the free stacks
  unsigned long long* heads = nullptr;  // MAPAD_GPOOL_SHARDS Treiber heads, 128 bytes apart
  uint64_t n_chunks = 0;
  int refs = 0;
};
std::mutex g_arena_mu;
DeviceArena g_arena[64];
}  // namespace

struct mapad_gpu {
  int device = 0;
  DeviceArena* arena = nullptr;
  int n_sm = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  IndexMeta meta;
  uint8_t* d_blob = nullptr;
  bool own_blob = false;
  mapad_params params;
  // resident batch
  uint64_t n_reads = 0, total_bases = 0;
  BatchPrep prep;
  bool have_batch = false;
  DevBuf<uint8_t> d_seq, d_qual;
  DevBuf<uint64_t> d_offsets;
  DevBuf<uint32_t> d_seeds;
  DevBuf<int16_t> d_starts;
  DevBuf<float> d_custom, d_bound, d_qualtab, d_dpen, d_dcomp;
  DevBuf<PenRow> d_delta;
  DevBuf<uint32_t> d_dsteps, d_deferred_a, d_deferred_b;
  DevBuf<ReadMid> d_mid;
  DevBuf<Cursors> d_cur;
  DevBuf<uint32_t> d_pool_tables;    // chunk tables of the groups of one launch
  DevBuf<uint32_t> d_order;          // read ids, longest first (work list of the group kernel)
  PinBuf<uint32_t> h_order;
  DevBuf<HitTmp> d_pool_hits;
  DevBuf<mapad_hit> d_hits;
  DevBuf<mapad_edit_op> d_ops;
  DevBuf<uint32_t> d_cigar;
  DevBuf<char> d_text;
  DevBuf<mapad_record> d_records;
  bool has_seeds = false;
  // pinned staging
  PinBuf<uint8_t> h_seq, h_qual;
  PinBuf<uint64_t> h_offsets;
  PinBuf<uint32_t> h_seeds;
  PinBuf<mapad_record> h_records;
  PinBuf<mapad_hit> h_hits;
  PinBuf<mapad_edit_op> h_ops;
  PinBuf<uint32_t> h_cigar;
  PinBuf<char> h_text;
  PinBuf<Cursors> h_cur;
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_wait = nullptr;  // blocking-sync event: a host thread waiting for a search launch (seconds to minutes) sleeps
  size_t ws_budget = 0;
};

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                 \
      return e_ == cudaErrorMemoryAllocation ? MAPAD_ENOMEM : MAPAD_ECUDA;                         \
    }                                                                                              \
  } while (0)

static int pick_device(int device, std::string& err) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) { err = "no CUDA device available (this library has no CPU fallback)"; return MAPAD_ENODEV; }
  if (device < 0 || device >= n) { err = "device index out of range"; return MAPAD_EINVAL; }
  return MAPAD_OK;
}

// Waits for the handle's stream without spinning: with one host thread per chunk in flight (and one process per GPU) the
// default spin-wait of cudaStreamSynchronize would keep tens of host cores busy for the length of every search launch.
static cudaError_t wait_stream(mapad_gpu* h) {
  cudaError_t e = cudaEventRecord(h->ev_wait, h->stream);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(h->ev_wait);
}

static int init_handle(mapad_gpu* h, int device) {
  h->device = device;
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  h->n_sm = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  h->own_stream = true;
  for (auto& e : h->ev) CK(cudaEventCreate(&e));
  CK(cudaEventCreateWithFlags(&h->ev_wait, cudaEventBlockingSync | cudaEventDisableTiming));
  return MAPAD_OK;
}

// The search workspace (chunk pool) is allocated after the first index blob of the device.  Size: MAPAD_WS_BYTES, else 75 %
// of the free device memory, leaving at least 1 GiB per handle announced with mapad_gpu_plan_handles (their batch buffers).
static int g_planned_handles[64] = {0};  // per device: handles still to be created
static int g_concurrency[64] = {0};      // per device: handles announced, i.e. launches expected in flight at once
static int alloc_workspace(mapad_gpu* h) {
  std::lock_guard<std::mutex> lock(g_arena_mu);
  DeviceArena& ar = g_arena[h->device & 63];
  if (ar.refs == 0) {
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const char* env = getenv("MAPAD_WS_BYTES");
    size_t budget = env ? (size_t)strtoull(env, nullptr, 10) : (size_t)(free_b * 0.75);
    if (!env) {
      const size_t reserve = (size_t)g_planned_handles[h->device & 63] << 30;
      if (budget + reserve > (size_t)(free_b * 0.9)) budget = (size_t)(free_b * 0.9) > reserve ? (size_t)(free_b * 0.9) - reserve : 0;
    }
    budget = std::max<size_t>(budget, (size_t)64 << 20);
    if (budget + ((size_t)128 << 20) > free_b) {
      h->err = "not enough free device memory for the search workspace (lower MAPAD_WS_BYTES)";
      return MAPAD_ENOMEM;
    }
    uint64_t n_chunks = budget >> MAPAD_GCHUNK_SHIFT;
    if (const char* e = getenv("MAPAD_TEST_POOL_CHUNKS")) n_chunks = std::min<uint64_t>(n_chunks, strtoull(e, nullptr, 10));  // test hook: tiny pool
    CK(cudaMalloc(&ar.base, (n_chunks << MAPAD_GCHUNK_SHIFT) + 4096));
    if (cudaMalloc(&ar.next, (n_chunks + 2) * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(&ar.heads, (size_t)MAPAD_GPOOL_SHARDS * MAPAD_GPOOL_HEAD_STRIDE * sizeof(unsigned long long)) != cudaSuccess) {
      cudaFree(ar.base); cudaFree(ar.next); ar.base = nullptr; ar.next = nullptr; h->err = "cudaMalloc(pool links)"; return MAPAD_ENOMEM;
    }
    ar.n_chunks = n_chunks;
    GChunkPool pool;
    pool.base = ar.base; pool.n_chunks = (uint32_t)n_chunks; pool.next = ar.next; pool.heads = ar.heads;
    k_gpool_init<<<(unsigned)((std::max<uint64_t>(n_chunks, MAPAD_GPOOL_SHARDS) + 255) / 256), 256>>>(pool);
    CK(cudaDeviceSynchronize());
  }
  ar.refs += 1;
  h->arena = &ar;
  h->ws_budget = (size_t)ar.n_chunks << MAPAD_GCHUNK_SHIFT;
  return MAPAD_OK;
}
static void release_workspace(mapad_gpu* h) {
  if (!h->arena) return;
  std::lock_guard<std::mutex> lock(g_arena_mu);
  DeviceArena& ar = *h->arena;
  h->arena = nullptr;
  if (--ar.refs == 0) {
    cudaFree(ar.base); cudaFree(ar.next); cudaFree(ar.heads);
    ar.base = nullptr; ar.next = nullptr; ar.heads = nullptr; ar.n_chunks = 0;
  }
}

extern "C" {

int mapad_gpu_create(const mapad_index* index, const mapad_params* params, int device, mapad_gpu** out) {
  if (!index || !params || !out) return MAPAD_EINVAL;
  *out = nullptr;
  std::string err;
  int rc = pick_device(device, err);
  if (rc) { fprintf(stderr, "mapad_gpu_create: %s\n", err.c_str()); return rc; }
  mapad_gpu* h = new (std::nothrow) mapad_gpu();
  if (!h) return MAPAD_ENOMEM;
  rc = init_handle(h, device);
  if (rc) { fprintf(stderr, "mapad_gpu_create: %s\n", h->err.c_str()); mapad_gpu_destroy(h); return rc; }
  h->params = *params;
  std::vector<uint8_t> blob;
  const char* fw = getenv("MAPAD_FORCE_WIDE");
  rc = build_device_blob(*reinterpret_cast<const HostIndex*>(index), h->meta, blob, fw && fw[0] == '1' ? 1 : -1);
  if (rc) { mapad_gpu_destroy(h); return rc; }
  cudaError_t e = cudaMalloc(&h->d_blob, blob.size());
  if (e != cudaSuccess) { mapad_gpu_destroy(h); return MAPAD_ENOMEM; }
  h->own_blob = true;
  e = cudaMemcpy(h->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { mapad_gpu_destroy(h); return MAPAD_ECUDA; }
  rc = alloc_workspace(h);
  if (rc) { fprintf(stderr, "mapad_gpu_create: %s\n", h->err.c_str()); mapad_gpu_destroy(h); return rc; }
  *out = h;
  return MAPAD_OK;
}

uint64_t mapad_gpu_index_meta_size(void) { return sizeof(IndexMeta); }

int mapad_gpu_export_index(mapad_gpu* h, void* meta_out, void** dev_ptr_out, uint64_t* dev_bytes_out) {
  if (!h || !meta_out || !dev_ptr_out || !dev_bytes_out) return MAPAD_EINVAL;
  memcpy(meta_out, &h->meta, sizeof(IndexMeta));
  *dev_ptr_out = h->d_blob;
  *dev_bytes_out = h->meta.total_bytes;
  return MAPAD_OK;
}

int mapad_gpu_copy_index_to(mapad_gpu* h, void* dst_dev_ptr, uint64_t dst_bytes) {
  if (!h || !dst_dev_ptr || dst_bytes < h->meta.total_bytes) return MAPAD_EINVAL;
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpy(dst_dev_ptr, h->d_blob, h->meta.total_bytes, cudaMemcpyDeviceToDevice));
  return MAPAD_OK;
}

int mapad_gpu_create_from_device_blob(const void* meta, void* dev_ptr, uint64_t dev_bytes, int take_ownership,
                                      const mapad_index* /*contigs_and_symbols*/, const mapad_params* params, int device,
                                      mapad_gpu** out) {
  if (!meta || !dev_ptr || !params || !out) return MAPAD_EINVAL;
  *out = nullptr;
  std::string err;
  int rc = pick_device(device, err);
  if (rc) return rc;
  mapad_gpu* h = new (std::nothrow) mapad_gpu();
  if (!h) return MAPAD_ENOMEM;
  rc = init_handle(h, device);
  if (rc) { mapad_gpu_destroy(h); return rc; }
  memcpy(&h->meta, meta, sizeof(IndexMeta));
  if (h->meta.total_bytes != dev_bytes) { mapad_gpu_destroy(h); return MAPAD_EINDEX; }
  h->params = *params;
  h->d_blob = (uint8_t*)dev_ptr;
  h->own_blob = take_ownership != 0;
  rc = alloc_workspace(h);
  if (rc) { h->own_blob = false; mapad_gpu_destroy(h); return rc; }
  *out = h;
  return MAPAD_OK;
}

int mapad_gpu_clone_to_device(mapad_gpu* src, int device, mapad_gpu** out) {
  if (!src || !out) return MAPAD_EINVAL;
  *out = nullptr;
  std::string err;
  int rc = pick_device(device, err);
  if (rc) return rc;
  mapad_gpu* h = new (std::nothrow) mapad_gpu();
  if (!h) return MAPAD_ENOMEM;
  rc = init_handle(h, device);
  if (rc) { src->err = h->err; mapad_gpu_destroy(h); return rc; }
  h->meta = src->meta;
  h->params = src->params;
  if (device == src->device) {  // same GPU: share the resident blob
    h->d_blob = src->d_blob;
    h->own_blob = false;
  } else {  // another GPU of the box: one peer copy of the re-laid-out blob (NVLink when peer access exists)
    if (cudaMalloc(&h->d_blob, h->meta.total_bytes) != cudaSuccess) { mapad_gpu_destroy(h); return MAPAD_ENOMEM; }
    h->own_blob = true;
    if (cudaMemcpyPeer(h->d_blob, device, src->d_blob, src->device, h->meta.total_bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
      mapad_gpu_destroy(h);
      return MAPAD_ECUDA;
    }
  }
  rc = alloc_workspace(h);
  if (rc) { src->err = h->err; mapad_gpu_destroy(h); return rc; }
  *out = h;
  return MAPAD_OK;
}

int mapad_gpu_plan_handles(int device, int n_handles) {
  if (device < 0 || device >= 64) return MAPAD_EINVAL;
  g_planned_handles[device] = n_handles > 0 ? n_handles : 0;
  g_concurrency[device] = n_handles > 0 ? n_handles : 0;
  return MAPAD_OK;
}

int mapad_gpu_set_params(mapad_gpu* h, const mapad_params* params) {
  if (!h || !params) return MAPAD_EINVAL;
  h->params = *params;
  return MAPAD_OK;
}

int mapad_gpu_set_stream(mapad_gpu* h, void* cuda_stream) {
  if (!h) return MAPAD_EINVAL;
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  if (cuda_stream) { h->stream = (cudaStream_t)cuda_stream; h->own_stream = false; }
  else { if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return MAPAD_ECUDA; h->own_stream = true; }
  return MAPAD_OK;
}

const char* mapad_gpu_last_error(const mapad_gpu* h) { return h ? h->err.c_str() : "null handle"; }

void mapad_gpu_destroy(mapad_gpu* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->own_blob && h->d_blob) cudaFree(h->d_blob);
  release_workspace(h);
  h->d_seq.release(); h->d_qual.release(); h->d_offsets.release(); h->d_seeds.release(); h->d_starts.release();
  h->d_custom.release(); h->d_bound.release(); h->d_qualtab.release(); h->d_dpen.release(); h->d_dcomp.release();
  h->d_delta.release(); h->d_dsteps.release(); h->d_deferred_a.release(); h->d_deferred_b.release(); h->d_mid.release();
  h->d_cur.release(); h->d_pool_tables.release(); h->d_pool_hits.release(); h->d_order.release(); h->h_order.release(); h->d_hits.release(); h->d_ops.release(); h->d_cigar.release(); h->d_text.release();
  h->d_records.release();
  h->h_seq.release(); h->h_qual.release(); h->h_offsets.release(); h->h_seeds.release(); h->h_records.release();
  h->h_hits.release(); h->h_ops.release(); h->h_cigar.release(); h->h_text.release(); h->h_cur.release();
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  if (h->ev_wait) cudaEventDestroy(h->ev_wait);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

}  // extern "C"

// ---- batch pipeline ---------------------------------------------------------------------------------
static int upload_batch(mapad_gpu* h, const mapad_reads* in) {
  int rc = prepare_batch(h->params, *in, h->prep);
  if (rc) { h->err = "invalid read batch"; return rc; }
  const uint64_t n = in->n_reads, tb = h->prep.total_bases;
  const uint64_t base0 = n ? in->offsets[0] : 0;
  h->n_reads = n; h->total_bases = tb;
  CK(h->h_seq.reserve(tb + 1)); CK(h->h_qual.reserve(tb + 1)); CK(h->h_offsets.reserve(n + 1));
  CK(h->d_seq.reserve(tb + 1)); CK(h->d_qual.reserve(tb + 1)); CK(h->d_offsets.reserve(n + 1));
  if (tb) { memcpy(h->h_seq.p, in->seq + base0, tb); memcpy(h->h_qual.p, in->qual + base0, tb); }
  for (uint64_t r = 0; r <= n; ++r) h->h_offsets.p[r] = n ? in->offsets[r] - base0 : 0;
  CK(cudaMemcpyAsync(h->d_seq.p, h->h_seq.p, tb, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_qual.p, h->h_qual.p, tb, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_offsets.p, h->h_offsets.p, (n + 1) * 8, cudaMemcpyHostToDevice, h->stream));
  {  // work list of the search kernel: longest reads first (counting sort by length; work per read grows steeply with it)
    CK(h->h_order.reserve(n + 1)); CK(h->d_order.reserve(n + 1));
    const uint32_t max_len = h->prep.max_len;
    std::vector<uint32_t> cnt((size_t)max_len + 2, 0);
    for (uint64_t r = 0; r < n; ++r) cnt[max_len - (uint32_t)(in->offsets[r + 1] - in->offsets[r]) + 1] += 1;
    for (size_t i = 1; i < cnt.size(); ++i) cnt[i] += cnt[i - 1];
    for (uint64_t r = 0; r < n; ++r) h->h_order.p[cnt[max_len - (uint32_t)(in->offsets[r + 1] - in->offsets[r])]++] = (uint32_t)r;
    CK(cudaMemcpyAsync(h->d_order.p, h->h_order.p, n * 4, cudaMemcpyHostToDevice, h->stream));
  }
  h->has_seeds = in->seeds != nullptr;
  if (h->has_seeds) {
    CK(h->h_seeds.reserve(n + 1)); CK(h->d_seeds.reserve(n + 1));
    memcpy(h->h_seeds.p, in->seeds, n * 4);
    CK(cudaMemcpyAsync(h->d_seeds.p, h->h_seeds.p, n * 4, cudaMemcpyHostToDevice, h->stream));
  }
  if (!h->prep.starts.empty()) {
    CK(h->d_starts.reserve(n + 1));
    CK(cudaMemcpyAsync(h->d_starts.p, h->prep.starts.data(), n * 2, cudaMemcpyHostToDevice, h->stream));
  }
  if (h->prep.dp.model == MODEL_TABLE) {
    CK(h->d_custom.reserve(4 * tb + 4));
    const float* src = in->custom_penalties ? in->custom_penalties + 4 * base0 : h->prep.custom_pen.data();
    CK(cudaMemcpyAsync(h->d_custom.p, src, 4 * tb * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  }
  CK(h->d_bound.reserve(h->prep.bound_table.size()));
  CK(cudaMemcpyAsync(h->d_bound.p, h->prep.bound_table.data(), h->prep.bound_table.size() * 4, cudaMemcpyHostToDevice, h->stream));
  CK(h->d_qualtab.reserve(256));
  CK(cudaMemcpyAsync(h->d_qualtab.p, h->prep.qual_table, 256 * 4, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));  // prep.* host vectors may be reused afterwards
  h->have_batch = true;
  return MAPAD_OK;
}


// ---- K2 driver: the group kernel (search_group.cuh) ---------------------------------------------
// One launch maps every read of the batch; a read is only handed back (deferred) when the chunk pool runs dry, and
// then re-run with fewer groups in flight, i.e. more pool per group.
#ifndef MAPAD_PREFETCH_WIDE_DEFAULT
#define MAPAD_PREFETCH_WIDE_DEFAULT 0u
#endif
struct GroupShape { int g, topl, minb; };
static GroupShape group_shape(bool wide) {  // tuning knobs, read per batch: MAPAD_GROUP = lanes per read, MAPAD_TOPL = heap lines in shared memory
  // defaults (profiles/r2_summary.md): small references (narrow layout, light reads) are issue-bound -> 4 lanes per read;
  // hg19-scale references (wide layout, reads of 1e5 - 1e7 frames) are bound by pool capacity and by the latency of
  // their heaviest reads -> one read per warp
  GroupShape s{wide ? 32 : 4, 11, MAPAD_GROUP_MIN_BLOCKS};
  if (const char* e = getenv("MAPAD_GROUP")) s.g = atoi(e);
  if (s.g != 1 && s.g != 4 && s.g != 8 && s.g != 16 && s.g != 32) s.g = wide ? 32 : 4;
  s.topl = s.g == 1 ? 3 : (s.g >= 16 ? 43 : 11);
  // one read per warp only: kernels compiled for 20 / 24 resident warps per SM (96 / 80 registers), MAPAD_GROUPS_PER_SM selects
  if (const char* e = getenv("MAPAD_GROUPS_PER_SM")) { const int m = atoi(e); if (s.g == 32 && (m == 20 || m == 24)) s.minb = m; }
  if (const char* e = getenv("MAPAD_TOPL")) { const int t = atoi(e); if (t == 3 || t == 11 || t == 43 || t == 171) s.topl = t; }
  // instantiated combinations: G = 1: 3 | 11, G = 4: 11, G = 8: 11 | 43, G = 16: 43, G = 32: 11 | 43 | 171 (10.7 KiB per read)
  if (s.g == 1 && s.topl > 11) s.topl = 11;
  if (s.g == 4) s.topl = 11;
  if (s.g == 8 && s.topl != 43) s.topl = 11;
  if (s.g == 16) s.topl = 43;
  if (s.g == 32 && s.topl == 3) s.topl = 11;
  return s;
}

// Latency hiding for deep heaps (GroupLaunch::prefetch): bit 0 = next family lines of a trickle-down, bit 1 = occ blocks of
// the popped frame.  MAPAD_TRICKLE_PREFETCH=<0..3> overrides the default (profiles/r2_summary.md).
static uint32_t prefetch_mode(bool wide) {
  if (const char* e = getenv("MAPAD_TRICKLE_PREFETCH")) return (uint32_t)atoi(e) & 3u;
  return wide ? MAPAD_PREFETCH_WIDE_DEFAULT : 0u;
}

template <bool WIDE, int G, int TOPL, int MINB = MAPAD_GROUP_MIN_BLOCKS>
static cudaError_t launch_group_kernel(const GroupLaunch<WIDE>& a, uint32_t n_groups, cudaStream_t stream) {
  constexpr int gpb = MAPAD_GROUP_BLOCK / G;
  const size_t smem = (size_t)gpb * TOPL * 64;
  cudaError_t e = cudaFuncSetAttribute(k_search_group<WIDE, G, TOPL, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_search_group<WIDE, G, TOPL, MINB><<<n_groups / gpb, MAPAD_GROUP_BLOCK, smem, stream>>>(a);
  return cudaGetLastError();
}
template <bool WIDE>
static cudaError_t launch_group_dispatch(const GroupShape& sh, const GroupLaunch<WIDE>& a, uint32_t n_groups, cudaStream_t stream) {
#define MAPAD_GROUP_CASE(G_, T_) if (sh.g == G_ && sh.topl == T_) return launch_group_kernel<WIDE, G_, T_>(a, n_groups, stream)
  MAPAD_GROUP_CASE(1, 3); MAPAD_GROUP_CASE(1, 11);
  MAPAD_GROUP_CASE(4, 11);
  MAPAD_GROUP_CASE(8, 11); MAPAD_GROUP_CASE(8, 43);
  MAPAD_GROUP_CASE(16, 43);
  if (sh.g == 32 && sh.topl == 43 && sh.minb == 20) return launch_group_kernel<WIDE, 32, 43, 20>(a, n_groups, stream);
  if (sh.g == 32 && sh.topl == 43 && sh.minb == 24) return launch_group_kernel<WIDE, 32, 43, 24>(a, n_groups, stream);
  MAPAD_GROUP_CASE(32, 11); MAPAD_GROUP_CASE(32, 43); MAPAD_GROUP_CASE(32, 171);
#undef MAPAD_GROUP_CASE
  return cudaErrorInvalidConfiguration;  // group_shape() only produces the combinations above
}

template <bool WIDE>
static int search_with_groups(mapad_gpu* h, const DevIndex& ix, const DevParams& P, const ReadBatch& rb, uint64_t& launches,
                              const std::function<void(const char*, uint64_t, uint64_t, uint64_t)>& trace) {
  const uint64_t n = h->n_reads;
  GroupShape sh = group_shape(WIDE);
  if (sh.topl == 3 && sh.g != 1) sh.topl = 11;
  if (sh.topl != 43) sh.minb = MAPAD_GROUP_MIN_BLOCKS;
  const uint32_t gpb = (uint32_t)(MAPAD_GROUP_BLOCK / sh.g);
  GroupLaunch<WIDE> a;
  a.ix = ix; a.P = P; a.rb = rb;
  a.bound_table = h->d_bound.p; a.delta = h->d_delta.p; a.dcomp = h->d_dcomp.p;
  if ((uint64_t)P.edit_tree_limit + 64 > 0x7fffffffull || (uint64_t)P.stack_limit + 64 > 0x7fffffffull) {
    h->err = "stack / edit-tree limits above 2^31 are not supported";
    return MAPAD_EINVAL;
  }
  a.max_nodes = P.edit_tree_limit + 64u;
  a.max_heap = P.stack_limit + 64u;
  a.nt = (a.max_nodes >> (MAPAD_GCHUNK_SHIFT - 5u)) + 1u;
  a.ht = (heap_lines_for(a.max_heap) >> (MAPAD_GCHUNK_SHIFT - 6u)) + 1u;
  const DeviceArena& ar = *h->arena;
  const uint64_t n_chunks = ar.n_chunks;
  // resident groups per SM: 16 warps (G = 8: 64 reads per SM); per-thread mode: 512 threads
  uint64_t per_sm = sh.g == 1 ? 512 : (sh.g == 32 ? 16 : (uint64_t)(512 / sh.g));
  if (const char* e = getenv("MAPAD_GROUPS_PER_SM")) per_sm = std::max<uint64_t>(1, strtoull(e, nullptr, 10));
  uint64_t slots = per_sm * (uint64_t)h->n_sm;
  // Launch share.  With C launches announced to be in flight at once (mapad_gpu_plan_handles) a launch takes 2/C of the
  // resident capacity, so that all chunks advance together and the heaviest reads of EVERY chunk start at once (longest-
  // first order) instead of waiting for the block slots of the launches enqueued before them.  This only works while the
  // reads in flight fit the pool: with G = 8 (9 472 reads in flight) on hg19-scale chunks it put the heavy reads of all
  // chunks in flight at once, the pool ran dry and the run got 2-3x slower (profiles/r2_summary.md) — there one read per
  // warp (G = 32, 2 368 reads in flight) is the setting that keeps the footprint inside the pool.
  // MAPAD_LAUNCH_SHARE=<divisor> overrides (1 = every launch sized for the whole GPU).
  {
    const int conc = g_concurrency[h->device & 63];
    uint64_t d = conc > 2 ? (uint64_t)conc / 2 : 1;
    if (const char* e = getenv("MAPAD_LAUNCH_SHARE")) d = std::max<uint64_t>(1, strtoull(e, nullptr, 10));
    slots = std::max<uint64_t>(4ull * gpb, slots / d);
  }
  if (const char* e = getenv("MAPAD_GROUPS")) slots = std::max<uint64_t>(1, strtoull(e, nullptr, 10));
  slots = std::min<uint64_t>(slots, n_chunks / 4);  // two base chunks per group, at least half of the pool for growth
  const uint32_t profile_iters = getenv("MAPAD_PROFILE_ITERS") ? (uint32_t)strtoul(getenv("MAPAD_PROFILE_ITERS"), nullptr, 0) : 0u;
  uint32_t n_work = (uint32_t)n;
  const uint32_t* work = h->d_order.p;
  uint32_t* deferred = h->d_deferred_a.p;
  // A read is handed back (deferred) when the shared pool stays dry for longer than the read's patience.  Deferred reads are
  // re-run with fewer groups of this handle in flight; the last resort is one read per launch that waits for the pool.
  bool serial = false;
  for (int attempt = 0; n_work > 0; ++attempt) {
    uint64_t use = std::min<uint64_t>(slots, ((uint64_t)n_work + gpb - 1) / gpb * gpb);
    use = use / gpb * gpb;
    if (use < gpb) {
      if (n_chunks < 2ull * gpb + 2) { h->err = "search workspace does not fit the device memory budget"; return MAPAD_ELIMIT; }
      use = gpb;
    }
    CK(h->d_pool_tables.reserve(use * (size_t)(a.nt + a.ht)));
    CK(h->d_pool_hits.reserve(use * MAPAD_MAX_HITS));
    a.pool.base = ar.base;
    a.pool.n_chunks = (uint32_t)n_chunks;
    a.pool.next = ar.next;
    a.pool.heads = ar.heads;
    a.tables = h->d_pool_tables.p;
    a.hit_base = h->d_pool_hits.p;
    a.cur = h->d_cur.p; a.mid = h->d_mid.p;
    a.hit_pool = h->d_hits.p; a.hit_cap = (uint32_t)std::min<size_t>(h->d_hits.cap, 0xffffffffu);
    a.op_pool = h->d_ops.p; a.op_cap = (uint32_t)std::min<size_t>(h->d_ops.cap, 0xffffffffu);
    a.iter_budget = profile_iters;
    a.flags_or = attempt ? 2u : 0u;
    a.patient = serial ? 1u : 0u;
    a.prefetch = prefetch_mode(WIDE);
    a.deferred_list = deferred;
    uint32_t n_def = 0;
    const uint32_t n_launches = serial ? n_work : 1u;
    for (uint32_t l = 0; l < n_launches; ++l) {
      a.work_list = serial ? work + l : work;
      a.n_work = serial ? 1u : n_work;
      CK(cudaMemsetAsync(h->d_cur.p, 0, 2 * sizeof(uint32_t), h->stream));  // queue head + deferred counter
      CK(launch_group_dispatch<WIDE>(sh, a, (uint32_t)(serial ? gpb : use), h->stream));
      ++launches;
      CK(cudaMemcpyAsync(h->h_cur.p, h->d_cur.p, sizeof(Cursors), cudaMemcpyDeviceToHost, h->stream));
      CK(wait_stream(h));
      CK(cudaGetLastError());
      if (profile_iters) { h->err = "MAPAD_PROFILE_ITERS is set: the search was cut short for profiling, no results"; return MAPAD_ELIMIT; }
      if (h->h_cur.p->overflow & MAPAD_POOL_TIMEOUT_FLAG) { h->err = "the device-wide chunk pool stayed empty for 20 s (workspace too small for the reads in flight)"; return MAPAD_ELIMIT; }
      n_def = h->h_cur.p->n_deferred;
      if (serial && n_def) { h->err = "a read exceeded the search workspace even with the whole pool to itself"; return MAPAD_ELIMIT; }
    }
    trace(serial ? "group (one read per launch)" : "group", n_work, n_def, use);
    if (n_def == 0) break;
    if (use <= gpb && n_def >= n_work) serial = true;  // no progress with the smallest grid: the reads block each other
    work = deferred;
    deferred = deferred == h->d_deferred_a.p ? h->d_deferred_b.p : h->d_deferred_a.p;
    n_work = n_def;
    slots = std::max<uint64_t>(gpb, use / 8);
  }
  return MAPAD_OK;
}

template <bool WIDE>
static int run_batch(mapad_gpu* h, uint32_t flags, mapad_results* out) {
  const uint64_t n = h->n_reads, tb = h->total_bases;
  const DevParams& P = h->prep.dp;
  uint64_t launches = 0;
  // MAPAD_TRACE=1: host-clock timeline of the lanes of every batch on stderr (tuning aid)
  static const int trace_level = getenv("MAPAD_TRACE") ? atoi(getenv("MAPAD_TRACE")) : 0;  // 1: timeline, 2: + lane utilisation
  static const bool trace_on = trace_level > 0;
  static const auto trace_epoch = std::chrono::steady_clock::now();
  auto now_s = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - trace_epoch).count(); };
  double trace_t0 = trace_on ? now_s() : 0.0;
  auto trace = [&](const char* what, uint64_t n_in, uint64_t n_out, uint64_t cap_) {
    if (!trace_on) return;
    const double t1 = now_s();
    fprintf(stderr, "[mapad trace] handle=%p %s start=%.3f end=%.3f reads=%llu deferred=%llu cap=%llu\n", (void*)h, what, trace_t0, t1,
            (unsigned long long)n_in, (unsigned long long)n_out, (unsigned long long)cap_);
    trace_t0 = t1;
  };
  DevIndex ix{h->meta, h->d_blob};
  ReadBatch rb;
  rb.n_reads = n; rb.seq = h->d_seq.p; rb.qual = h->d_qual.p; rb.offsets = h->d_offsets.p;
  rb.seeds = h->has_seeds ? h->d_seeds.p : nullptr;
  rb.starts = h->prep.starts.empty() ? nullptr : h->d_starts.p;
  rb.custom_pen = P.model == MODEL_TABLE ? h->d_custom.p : nullptr;
  CK(h->d_delta.reserve(tb + 1)); CK(h->d_dpen.reserve(tb + 1)); CK(h->d_dcomp.reserve(tb + 1));
  CK(h->d_dsteps.reserve(n + 1)); CK(h->d_mid.reserve(n + 1)); CK(h->d_cur.reserve(1)); CK(h->h_cur.reserve(1));
  CK(h->d_deferred_a.reserve(n + 1)); CK(h->d_deferred_b.reserve(n + 1));
  CK(h->d_records.reserve(n + 1));
  // output pools (grown and the batch re-run if a cursor overshoots)
  size_t hit_cap = std::max(h->d_hits.cap, (size_t)(4 * n + 1024));
  size_t op_cap = std::max(h->d_ops.cap, (size_t)(4 * tb + 64 * n + 1024));
  size_t cig_cap = std::max(h->d_cigar.cap, (size_t)(8 * n + 1024));
  size_t text_cap = std::max(h->d_text.cap, (size_t)(24 * n + 1024));
  CK(cudaEventRecord(h->ev[1], h->stream));
  if (n) {
    // ---- K1: prologue ----
    {
      const int block = 256;
      const uint64_t warps_needed = n;
      int grid = (int)std::min<uint64_t>((warps_needed * 32 + block - 1) / block, (uint64_t)h->n_sm * 32);
      k_penalties<<<grid, block, 0, h->stream>>>(P, rb, h->d_qualtab.p, h->d_delta.p, h->d_dpen.p);
      ++launches;
      grid = (int)std::min<uint64_t>((n * 16 + block - 1) / block, (uint64_t)h->n_sm * 32);
      k_darray<WIDE><<<grid, block, 0, h->stream>>>(ix, P, rb, h->d_dpen.p, h->d_dcomp.p, h->d_dsteps.p);
      ++launches;
    }
  }
  CK(cudaEventRecord(h->ev[2], h->stream));
  for (int attempt = 0;; ++attempt) {
    CK(h->d_hits.reserve(hit_cap)); CK(h->d_ops.reserve(op_cap)); CK(h->d_cigar.reserve(cig_cap)); CK(h->d_text.reserve(text_cap));
    CK(cudaMemsetAsync(h->d_cur.p, 0, sizeof(Cursors), h->stream));
    // ---- K2: search ----
    {
      const int rc = search_with_groups<WIDE>(h, ix, P, rb, launches, trace);
      if (rc) return rc;
    }
    CK(cudaEventRecord(h->ev[3], h->stream));
    // K2 bump-allocates hits and edit operations; when a cursor overshot its pool the batch is re-run with larger pools
    // BEFORE K3 would read the truncated spans (h_cur was copied after the last search launch)
    const bool k2_over = n && (h->h_cur.p->hit_cursor > h->d_hits.cap || h->h_cur.p->op_cursor > h->d_ops.cap || (h->h_cur.p->overflow & 1u));
    // ---- K3: epilogue ----
    if (n && !k2_over) {
      OutPools pools;
      pools.cigar = h->d_cigar.p; pools.cigar_cap = (uint32_t)std::min<size_t>(h->d_cigar.cap, 0xffffffffu);
      pools.cigar_cursor = &h->d_cur.p->cigar_cursor;
      pools.text = h->d_text.p; pools.text_cap = (uint32_t)std::min<size_t>(h->d_text.cap, 0xffffffffu);
      pools.text_cursor = &h->d_cur.p->text_cursor;
      pools.overflow = &h->d_cur.p->pad;
      const int block = 128;
      const int grid = (int)((n + block - 1) / block);
      k_epilogue<WIDE><<<grid, block, 0, h->stream>>>(ix, P, rb, h->d_bound.p, h->d_mid.p, h->d_dsteps.p, h->d_hits.p, h->d_ops.p, pools,
                                                      h->d_records.p);
      ++launches;
    }
    CK(cudaEventRecord(h->ev[4], h->stream));
    CK(cudaMemcpyAsync(h->h_cur.p, h->d_cur.p, sizeof(Cursors), cudaMemcpyDeviceToHost, h->stream));
    CK(wait_stream(h));
    CK(cudaGetLastError());
    trace("epilogue", n, 0, 0);
    const Cursors c = *h->h_cur.p;
    const bool over = k2_over || c.hit_cursor > h->d_hits.cap || c.op_cursor > h->d_ops.cap || c.cigar_cursor > h->d_cigar.cap ||
                      c.text_cursor > h->d_text.cap || (c.overflow & 1u) || c.pad;
    if (!over) break;
    if (attempt >= 3) { h->err = "output pools overflowed repeatedly"; return MAPAD_ELIMIT; }
    hit_cap = std::max<size_t>(hit_cap, (size_t)c.hit_cursor * 2 + 1024);
    op_cap = std::max<size_t>(op_cap, (size_t)c.op_cursor * 2 + 1024);
    cig_cap = std::max<size_t>(cig_cap, (size_t)c.cigar_cursor * 2 + 1024);
    text_cap = std::max<size_t>(text_cap, (size_t)c.text_cursor * 2 + 1024);
  }
  // ---- D2H ----
  const Cursors c = *h->h_cur.p;
  memset(out, 0, sizeof *out);
  out->n_reads = n;
  if (!(flags & MAPAD_BATCH_NO_D2H)) {
    CK(h->h_records.reserve(n + 1));
    CK(cudaMemcpyAsync(h->h_records.p, h->d_records.p, n * sizeof(mapad_record), cudaMemcpyDeviceToHost, h->stream));
    CK(h->h_cigar.reserve(c.cigar_cursor + 1)); CK(h->h_text.reserve(c.text_cursor + 1));
    CK(cudaMemcpyAsync(h->h_cigar.p, h->d_cigar.p, (size_t)c.cigar_cursor * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(h->h_text.p, h->d_text.p, c.text_cursor, cudaMemcpyDeviceToHost, h->stream));
    if (flags & MAPAD_BATCH_WANT_HITS) {
      CK(h->h_hits.reserve(c.hit_cursor + 1)); CK(h->h_ops.reserve(c.op_cursor + 1));
      CK(cudaMemcpyAsync(h->h_hits.p, h->d_hits.p, (size_t)c.hit_cursor * sizeof(mapad_hit), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaMemcpyAsync(h->h_ops.p, h->d_ops.p, (size_t)c.op_cursor * sizeof(mapad_edit_op), cudaMemcpyDeviceToHost, h->stream));
      out->hits = h->h_hits.p; out->n_hits = c.hit_cursor;
      out->edit_ops = h->h_ops.p; out->n_edit_ops = c.op_cursor;
    }
    out->records = h->h_records.p;
    out->cigar = h->h_cigar.p; out->n_cigar = c.cigar_cursor;
    out->text = h->h_text.p; out->n_text = c.text_cursor;
  }
  CK(cudaEventRecord(h->ev[5], h->stream));
  CK(wait_stream(h));
  cudaEventElapsedTime(&out->ms_h2d, h->ev[0], h->ev[1]);
  cudaEventElapsedTime(&out->ms_prologue, h->ev[1], h->ev[2]);
  cudaEventElapsedTime(&out->ms_search, h->ev[2], h->ev[3]);
  cudaEventElapsedTime(&out->ms_epilogue, h->ev[3], h->ev[4]);
  cudaEventElapsedTime(&out->ms_d2h, h->ev[4], h->ev[5]);
  cudaEventElapsedTime(&out->ms_total, h->ev[0], h->ev[5]);
  out->gpu_launches = launches;
  return MAPAD_OK;
}

extern "C" {

int mapad_gpu_map_batch(mapad_gpu* h, const mapad_reads* in, uint32_t flags, mapad_results* out) {
  if (!h || !out) return MAPAD_EINVAL;
  if (cudaSetDevice(h->device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return MAPAD_ECUDA; }
  CK(cudaEventRecord(h->ev[0], h->stream));
  if (!(flags & MAPAD_BATCH_RESIDENT)) {
    if (!in) return MAPAD_EINVAL;
    int rc = upload_batch(h, in);
    if (rc) return rc;
    if (flags & MAPAD_BATCH_UPLOAD_ONLY) {
      memset(out, 0, sizeof *out);
      out->n_reads = h->n_reads;
      return MAPAD_OK;
    }
  } else if (!h->have_batch) {
    h->err = "MAPAD_BATCH_RESIDENT without a previously uploaded batch";
    return MAPAD_EINVAL;
  }
  return h->meta.wide ? run_batch<true>(h, flags, out) : run_batch<false>(h, flags, out);
}

int mapad_gpu_gather_peak(int device, uint64_t table_bytes, uint32_t bytes_per_access, uint64_t n_accesses, double* gbps_out) {
  if (!gbps_out || (bytes_per_access != 16 && bytes_per_access != 32 && bytes_per_access != 64)) return MAPAD_EINVAL;
  std::string err;
  int rc = pick_device(device, err);
  if (rc) return rc;
  if (cudaSetDevice(device) != cudaSuccess) return MAPAD_ECUDA;
  uint4* table = nullptr;
  unsigned long long* sink = nullptr;
  table_bytes = std::max<uint64_t>(table_bytes / 64 * 64, 1 << 20);
  if (cudaMalloc(&table, table_bytes) != cudaSuccess) return MAPAD_ENOMEM;
  if (cudaMalloc(&sink, 8) != cudaSuccess) { cudaFree(table); return MAPAD_ENOMEM; }
  cudaMemset(table, 1, table_bytes);
  cudaMemset(sink, 0, 8);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  const int block = 256, grid = prop.multiProcessorCount * 8;
  const uint64_t threads = (uint64_t)block * grid;
  const uint64_t per_thread = std::max<uint64_t>(1, n_accesses / threads);
  const uint64_t n_units = table_bytes / bytes_per_access;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best_ms = 1e30f;
  for (int it = 0; it < 4; ++it) {
    cudaEventRecord(e0);
    if (bytes_per_access == 16) k_gather<1><<<grid, block>>>(table, n_units, per_thread, 1234 + it, sink);
    else if (bytes_per_access == 32) k_gather<2><<<grid, block>>>(table, n_units, per_thread, 1234 + it, sink);
    else k_gather<4><<<grid, block>>>(table, n_units, per_thread, 1234 + it, sink);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(table); cudaFree(sink); return MAPAD_ECUDA; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (it > 0 && ms < best_ms) best_ms = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(table); cudaFree(sink);
  *gbps_out = (double)per_thread * threads * bytes_per_access / (best_ms * 1e-3) / 1e9;
  return MAPAD_OK;
}

int mapad_gpu_debug_libm(int device, int fn, int iarg, uint64_t n, const float* in, float* out) {
  if (!in || !out || fn < 0 || fn > 3) return MAPAD_EINVAL;
  std::string err;
  int rc = pick_device(device, err);
  if (rc) return rc;
  if (cudaSetDevice(device) != cudaSuccess) return MAPAD_ECUDA;
  float *d_in = nullptr, *d_out = nullptr;
  if (cudaMalloc(&d_in, n * 4 + 4) != cudaSuccess) return MAPAD_ENOMEM;
  if (cudaMalloc(&d_out, n * 4 + 4) != cudaSuccess) { cudaFree(d_in); return MAPAD_ENOMEM; }
  cudaMemcpy(d_in, in, n * 4, cudaMemcpyHostToDevice);
  if (n) k_debug_libm<<<(unsigned)((n + 255) / 256), 256>>>(fn, iarg, n, d_in, d_out);
  cudaError_t e = cudaMemcpy(out, d_out, n * 4, cudaMemcpyDeviceToHost);
  cudaFree(d_in); cudaFree(d_out);
  return e == cudaSuccess ? MAPAD_OK : MAPAD_ECUDA;
}

int64_t mapad_format_xa(const mapad_index* index, const mapad_results* res, uint64_t read_idx, char* buf, uint64_t cap) {
  if (!index || !res || !buf || read_idx >= res->n_reads || !res->records) return MAPAD_EINVAL;
  const HostIndex* ix = reinterpret_cast<const HostIndex*>(index);
  const mapad_record& r = res->records[read_idx];
  std::string s;
  for (uint32_t a = 0; a < r.n_alts && a < 2; ++a) {  // mapping.rs:475-488
    const mapad_alt& al = r.alts[a];
    if (al.tid < 0 || (uint64_t)al.tid >= ix->contig_names.size()) return MAPAD_EINDEX;
    s += ix->contig_names[al.tid];
    s += al.strand ? ",-" : ",+";
    s += std::to_string(al.pos + 1);
    s += ",";
    for (uint32_t i = 0; i < al.cigar_len; ++i) {
      uint32_t v = res->cigar[al.cigar_off + i];
      s += std::to_string(v >> 4);
      s += "MID"[v & 15];
    }
    s += ",";
    s.append(res->text + al.md_off, al.md_len);
    s += "," + std::to_string(al.nm) + "," + std::to_string(al.interval_size) + ",";
    char num[64];
    snprintf(num, sizeof num, "%.2f", (double)al.alignment_score);  // Rust `{:.2}`
    s += num;
    s += ";";
  }
  if (s.size() + 1 > cap) return MAPAD_EINVAL;
  memcpy(buf, s.c_str(), s.size() + 1);
  return (int64_t)s.size();
}

}  // extern "C"
