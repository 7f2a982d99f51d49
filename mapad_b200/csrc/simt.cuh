// simt.cuh — lane-group primitives.  A "group" is G consecutive lanes of a warp (G = 1, 2, 4, 8, 16 or 32) that own one
// read together; every collective below involves exactly the lanes of the caller's group, so the groups of a warp may
// diverge from each other freely (independent thread scheduling, sm_70+).
//
// Device build: warp shuffles / votes with the group's member mask.  Plain C++ build (tests/emu): G == 1 is the
// identity; G > 1 calls hooks implemented by the SIMT emulator of the test-suite (tests/emu/simt_emu.cpp), which runs
// the lanes of a group as coroutines — the product library never contains that path.
#pragma once
#include <cstdint>

#include "dev_index.cuh"

#if !defined(__CUDA_ARCH__)
extern "C" uint32_t mapad_simt_emu_shfl(uint32_t v, int src_lane_in_group, int group_size);
extern "C" uint32_t mapad_simt_emu_ballot(int pred, int group_size);
extern "C" void mapad_simt_emu_sync(int group_size);
#endif

namespace mapad {

template <int G>
struct Grp {
  static_assert(G == 1 || G == 2 || G == 4 || G == 8 || G == 16 || G == 32, "group size");
#if defined(__CUDA_ARCH__)
  static __device__ __forceinline__ unsigned base() { return (threadIdx.x & 31u) & ~(unsigned)(G - 1); }
  static __device__ __forceinline__ unsigned mask() { return G >= 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << base()); }
  static __device__ __forceinline__ uint32_t shfl(uint32_t v, int src) { return G == 1 ? v : __shfl_sync(mask(), v, src, G); }
  static __device__ __forceinline__ uint32_t shfl_xor(uint32_t v, int m) { return G == 1 ? v : __shfl_xor_sync(mask(), v, m, G); }
  static __device__ __forceinline__ uint32_t ballot(bool p) {
    if (G == 1) return p ? 1u : 0u;
    const unsigned b = __ballot_sync(mask(), p);
    return G >= 32 ? b : ((b >> base()) & ((1u << (G & 31)) - 1u));
  }
  static __device__ __forceinline__ void sync() { if (G > 1) __syncwarp(mask()); }
#else
  static inline uint32_t shfl(uint32_t v, int src) { return G == 1 ? v : mapad_simt_emu_shfl(v, src, G); }
  static inline uint32_t ballot(bool p) { return G == 1 ? (p ? 1u : 0u) : mapad_simt_emu_ballot(p ? 1 : 0, G); }
  static inline void sync() { if (G > 1) mapad_simt_emu_sync(G); }
  // xor-shuffle on top of the indexed one; `self` is the caller's lane in the group
  static inline uint32_t shfl_xor_self(uint32_t v, int m, int self) { return G == 1 ? v : mapad_simt_emu_shfl(v, self ^ m, G); }
#endif
  static MAPAD_DEV float shfl_f(float v, int src) {
    union { float f; uint32_t u; } a, b;
    a.f = v; b.u = shfl(a.u, src);
    return b.f;
  }
  static MAPAD_DEV uint64_t shfl_u64(uint64_t v, int src) {
    const uint32_t lo = shfl((uint32_t)v, src), hi = shfl((uint32_t)(v >> 32), src);
    return (uint64_t)lo | ((uint64_t)hi << 32);
  }
};

// device-wide atomics used by the queues / bump allocators (plain operations in the single-threaded emulation)
MAPAD_DEV uint32_t dev_atomic_add(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
  return atomicAdd(p, v);
#else
  const uint32_t o = *p; *p = o + v; return o;
#endif
}
MAPAD_DEV void dev_atomic_or(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
  atomicOr(p, v);
#else
  *p |= v;
#endif
}

}  // namespace mapad
