// libm_emu.cuh — bit-exact restatements of the glibc 2.39 float functions the reference reaches
// through Rust's f32::{log2, exp2, log10} (src/map/sequence_difference_models.rs:205-206,
// src/map/mapping.rs:669,684,694), usable on the device.  glibc evaluates these in double
// precision with small tables (sysdeps/ieee754/flt-32/e_{log2f,exp2f,logf,log10f}.c); the
// tables below are those constants.  tests/test_libm_emu.py compares every function against the
// host's libm (exhaustively for the full float range when MAPAD_EXHAUSTIVE=1, sampled otherwise)
// and the GPU test does the same on the device.
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define MAPAD_HD __host__ __device__ __forceinline__
#define MAPAD_TAB __device__ __constant__
#else
#define MAPAD_HD inline
#define MAPAD_TAB static const
#endif

namespace mapad {
namespace emu {
#if defined(__CUDA_ARCH__)
#define MAPAD_EMU_TAB(x) d_##x
#else
#define MAPAD_EMU_TAB(x) h_##x
#endif
static const double h_LOG2F_TAB[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2},
    {0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2},
    {0x1.49539f0f010b0p+0, -0x1.7418b0a1fb77bp-2},
    {0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2},
    {0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2},
    {0x1.25e227b0b8ea0p+0, -0x1.97c1d1b3b7af0p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3},
    {0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4},
    {0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5},
    {0x1.0000000000000p+0, 0x0.0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4},
    {0x1.ca4b31f026aa0p-1, 0x1.476a9543891bap-3},
    {0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3},
    {0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2},
    {0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2},
    {0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2},
};
static const double h_LOGF_TAB[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2},
    {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2},
    {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3},
    {0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4},
    {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5},
    {0x1.0000000000000p+0, 0x0.0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},
    {0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},
    {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},
    {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2},
};
static const uint64_t h_EXP2F_TAB[32] = {
    0x3ff0000000000000ull,
    0x3fefd9b0d3158574ull,
    0x3fefb5586cf9890full,
    0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull,
    0x3fef54873168b9aaull,
    0x3fef387a6e756238ull,
    0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull,
    0x3feef1a7373aa9cbull,
    0x3feedea64c123422ull,
    0x3feece086061892dull,
    0x3feebfdad5362a27ull,
    0x3feeb42b569d4f82ull,
    0x3feeab07dd485429ull,
    0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull,
    0x3fee9f75e8ec5f74ull,
    0x3feea11473eb0187ull,
    0x3feea589994cce13ull,
    0x3feeace5422aa0dbull,
    0x3feeb737b0cdc5e5ull,
    0x3feec49182a3f090ull,
    0x3feed503b23e255dull,
    0x3feee89f995ad3adull,
    0x3feeff76f2fb5e47ull,
    0x3fef199bdd85529cull,
    0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull,
    0x3fef7c97337b9b5full,
    0x3fefa4afa2a490daull,
    0x3fefd0765b6e4540ull,
};
#if defined(__CUDACC__)
__device__ __constant__ double d_LOG2F_TAB[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2},
    {0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2},
    {0x1.49539f0f010b0p+0, -0x1.7418b0a1fb77bp-2},
    {0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2},
    {0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2},
    {0x1.25e227b0b8ea0p+0, -0x1.97c1d1b3b7af0p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3},
    {0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4},
    {0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5},
    {0x1.0000000000000p+0, 0x0.0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4},
    {0x1.ca4b31f026aa0p-1, 0x1.476a9543891bap-3},
    {0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3},
    {0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2},
    {0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2},
    {0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2},
};
__device__ __constant__ double d_LOGF_TAB[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2},
    {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2},
    {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3},
    {0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4},
    {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5},
    {0x1.0000000000000p+0, 0x0.0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},
    {0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},
    {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},
    {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2},
};
__device__ __constant__ uint64_t d_EXP2F_TAB[32] = {
    0x3ff0000000000000ull,
    0x3fefd9b0d3158574ull,
    0x3fefb5586cf9890full,
    0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull,
    0x3fef54873168b9aaull,
    0x3fef387a6e756238ull,
    0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull,
    0x3feef1a7373aa9cbull,
    0x3feedea64c123422ull,
    0x3feece086061892dull,
    0x3feebfdad5362a27ull,
    0x3feeb42b569d4f82ull,
    0x3feeab07dd485429ull,
    0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull,
    0x3fee9f75e8ec5f74ull,
    0x3feea11473eb0187ull,
    0x3feea589994cce13ull,
    0x3feeace5422aa0dbull,
    0x3feeb737b0cdc5e5ull,
    0x3feec49182a3f090ull,
    0x3feed503b23e255dull,
    0x3feee89f995ad3adull,
    0x3feeff76f2fb5e47ull,
    0x3fef199bdd85529cull,
    0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull,
    0x3fef7c97337b9b5full,
    0x3fefa4afa2a490daull,
    0x3fefd0765b6e4540ull,
};
#endif

MAPAD_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
MAPAD_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
MAPAD_HD uint64_t d2u(double f) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(f);
#else
  uint64_t u; memcpy(&u, &f, 8); return u;
#endif
}
MAPAD_HD double u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double f; memcpy(&f, &u, 8); return f;
#endif
}
// Explicitly rounded double ops: never contracted into FMAs, on either side.
MAPAD_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  volatile double r = a * b; return r;
#endif
}
MAPAD_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  volatile double r = a + b; return r;
#endif
}
MAPAD_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b; return r;
#endif
}
MAPAD_HD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b; return r;
#endif
}
MAPAD_HD float pos_inf() { return u2f(0x7f800000u); }
MAPAD_HD float quiet_nan() { return u2f(0x7fc00000u); }

// glibc __log2f
MAPAD_HD float log2f_glibc(float x) {
  const double A0 = -0x1.712b6f70a7e4dp-2, A1 = 0x1.ecabf496832e0p-2, A2 = -0x1.715479ffae3dep-1, A3 = 0x1.715475f35c8b8p+0;
  uint32_t ix = f2u(x);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix * 2 == 0) return -pos_inf();
    if (ix == 0x7f800000u) return x;
    if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return quiet_nan();
    ix = f2u(fmul(x, 0x1p23f));
    ix -= 23u << 23;
  }
  uint32_t tmp = ix - 0x3f330000u;
  int i = (int)((tmp >> (23 - 4)) % 16);
  uint32_t top = tmp & 0xff800000u;
  uint32_t iz = ix - top;
  int k = (int32_t)tmp >> 23;
  double invc = MAPAD_EMU_TAB(LOG2F_TAB)[i][0], logc = MAPAD_EMU_TAB(LOG2F_TAB)[i][1];
  double z = (double)u2f(iz);
  double r = dadd(dmul(z, invc), -1.0);
  double y0 = dadd(logc, (double)k);
  double r2 = dmul(r, r);
  double y = dadd(dmul(A1, r), A2);
  y = dadd(dmul(A0, r2), y);
  double p = dadd(dmul(A3, r), y0);
  y = dadd(dmul(y, r2), p);
  return (float)y;
}

// glibc __exp2f
MAPAD_HD float exp2f_glibc(float x) {
  if (x != x) return x;
  if (x >= 128.0f) return pos_inf();
  if (x <= -150.0f) return 0.0f;
  const double SHIFT = 0x1.8p+47;
  const double C0 = 0x1.c6af84b912394p-5, C1 = 0x1.ebfce50fac4f3p-3, C2 = 0x1.62e42ff0c52d6p-1;
  double xd = (double)x;
  double kd = dadd(xd, SHIFT);
  uint64_t ki = d2u(kd);
  kd = dadd(kd, -SHIFT);
  double r = dadd(xd, -kd);
  uint64_t t = MAPAD_EMU_TAB(EXP2F_TAB)[ki % 32];
  t += ki << (52 - 5);
  double s = u2d(t);
  double z = dadd(dmul(C0, r), C1);
  double r2 = dmul(r, r);
  double y = dadd(dmul(C2, r), 1.0);
  y = dadd(dmul(z, r2), y);
  y = dmul(y, s);
  return (float)y;
}

// glibc __logf
MAPAD_HD float logf_glibc(float x) {
  const double Ln2 = 0x1.62e42fefa39efp-1;
  const double A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
  uint32_t ix = f2u(x);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix * 2 == 0) return -pos_inf();
    if (ix == 0x7f800000u) return x;
    if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return quiet_nan();
    ix = f2u(fmul(x, 0x1p23f));
    ix -= 23u << 23;
  }
  uint32_t tmp = ix - 0x3f330000u;
  int i = (int)((tmp >> (23 - 4)) % 16);
  int k = (int32_t)tmp >> 23;
  uint32_t iz = ix - (tmp & 0xff800000u);
  double invc = MAPAD_EMU_TAB(LOGF_TAB)[i][0], logc = MAPAD_EMU_TAB(LOGF_TAB)[i][1];
  double z = (double)u2f(iz);
  double r = dadd(dmul(z, invc), -1.0);
  double y0 = dadd(logc, dmul((double)k, Ln2));
  double r2 = dmul(r, r);
  double y = dadd(dmul(A1, r), A2);
  y = dadd(dmul(A0, r2), y);
  y = dadd(dmul(y, r2), dadd(y0, r));
  return (float)y;
}

// glibc __ieee754_log10f (fdlibm-style float code on top of logf)
MAPAD_HD float log10f_glibc(float x) {
  const float two25 = 3.3554432000e+07f, ivln10 = 4.3429449201e-01f, log10_2hi = 3.0102920532e-01f,
              log10_2lo = 7.9034151668e-07f;
  int32_t hx = (int32_t)f2u(x);
  int32_t k = 0;
  if (hx < 0x00800000) {
    if ((hx & 0x7fffffff) == 0) return -pos_inf();
    if (hx < 0) return quiet_nan();
    k -= 25;
    x = fmul(x, two25);
    hx = (int32_t)f2u(x);
  }
  if (hx >= 0x7f800000) return fadd(x, x);
  k += (hx >> 23) - 127;
  int32_t i = (int32_t)(((uint32_t)k & 0x80000000u) >> 31);
  hx = (hx & 0x007fffff) | ((0x7f - i) << 23);
  float y = (float)(k + i);
  float lf = logf_glibc(u2f((uint32_t)hx));
  float z = fadd(fmul(y, log10_2lo), fmul(ivln10, lf));
  return fadd(z, fmul(y, log10_2hi));
}

// compiler-rt __powisf2 (Rust f32::powi)
MAPAD_HD float powi_rt(float a, int b) {
  const bool recip = b < 0;
  float r = 1.0f;
  while (true) {
    if (b & 1) r = fmul(r, a);
    b /= 2;
    if (b == 0) break;
    a = fmul(a, a);
  }
#if defined(__CUDA_ARCH__)
  return recip ? __fdiv_rn(1.0f, r) : r;
#else
  return recip ? 1.0f / r : r;
#endif
}

}  // namespace emu
}  // namespace mapad
