// common.cuh - scalar helpers shared by the sm_100a ensemble kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/diffrax_b200.h"

namespace dfx {

constexpr int kMaxDim = 8;
constexpr int kWarp = 32;
constexpr unsigned kFullMask = 0xffffffffu;

template <class R> struct Num;
template <> struct Num<double> {
  using uint_t = unsigned long long;
  using int_t = long long;
  static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
  static __device__ __forceinline__ double nan() { return __longlong_as_double(0x7ff8000000000000LL); }
  static __device__ __forceinline__ double eps() { return 2.220446049250313e-16; }
  static __device__ __forceinline__ long long bits(double x) { return __double_as_longlong(x); }
  static __device__ __forceinline__ double from_bits(long long b) { return __longlong_as_double(b); }
};
template <> struct Num<float> {
  using uint_t = unsigned int;
  using int_t = int;
  static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
  static __device__ __forceinline__ float nan() { return __int_as_float(0x7fc00000); }
  static __device__ __forceinline__ float eps() { return 1.1920929e-07f; }
  static __device__ __forceinline__ int bits(float x) { return __float_as_int(x); }
  static __device__ __forceinline__ float from_bits(int b) { return __int_as_float(b); }
};

__device__ __forceinline__ double r_abs(double x) { return fabs(x); }
__device__ __forceinline__ float r_abs(float x) { return fabsf(x); }
__device__ __forceinline__ double r_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float r_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double r_pow(double x, double y) { return pow(x, y); }
__device__ __forceinline__ float r_pow(float x, float y) { return powf(x, y); }
__device__ __forceinline__ double r_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float r_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double r_min(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float r_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double r_sin(double x) { return sin(x); }
__device__ __forceinline__ float r_sin(float x) { return sinf(x); }
__device__ __forceinline__ double r_exp(double x) { return exp(x); }
__device__ __forceinline__ float r_exp(float x) { return expf(x); }
__device__ __forceinline__ double r_log1p(double x) { return log1p(x); }
__device__ __forceinline__ float r_log1p(float x) { return log1pf(x); }
__device__ __forceinline__ double r_tanh(double x) { return tanh(x); }
__device__ __forceinline__ float r_tanh(float x) { return tanhf(x); }
__device__ __forceinline__ bool r_isnan(double x) { return x != x; }
__device__ __forceinline__ bool r_isnan(float x) { return x != x; }
__device__ __forceinline__ bool r_isinf(double x) { return fabs(x) == Num<double>::inf(); }
__device__ __forceinline__ bool r_isinf(float x) { return fabsf(x) == Num<float>::inf(); }

// jnp.maximum / jnp.minimum propagate NaN; CUDA fmax/fmin do not.
template <class R> __device__ __forceinline__ R jnp_max(R a, R b) { return (a != a || b != b) ? Num<R>::nan() : r_max(a, b); }
template <class R> __device__ __forceinline__ R jnp_min(R a, R b) { return (a != a || b != b) ? Num<R>::nan() : r_min(a, b); }

// eqxi.prevbefore applied n times == nextafter(x, -inf) n times (_integrate.py:320-322),
// done in one integer step on the ordered bit pattern.
template <class R> __device__ __host__ inline R prev_n(R x, int n);
template <> __device__ __host__ inline double prev_n<double>(double x, int n) {
  if (x != x) return x;
  long long b;
#ifdef __CUDA_ARCH__
  b = __double_as_longlong(x);
#else
  memcpy(&b, &x, 8);
#endif
  long long key = b >= 0 ? b : -(b & 0x7fffffffffffffffLL);
  if (b == 0x7ff0000000000000LL) { /* +inf -> largest finite, then n-1 more */ }
  key -= n;
  long long ob = key >= 0 ? key : (long long)(0x8000000000000000ULL | (unsigned long long)(-key));
  double out;
#ifdef __CUDA_ARCH__
  out = __longlong_as_double(ob);
#else
  memcpy(&out, &ob, 8);
#endif
  return out;
}
template <> __device__ __host__ inline float prev_n<float>(float x, int n) {
  if (x != x) return x;
  int b;
#ifdef __CUDA_ARCH__
  b = __float_as_int(x);
#else
  memcpy(&b, &x, 4);
#endif
  int key = b >= 0 ? b : -(b & 0x7fffffff);
  key -= n;
  int ob = key >= 0 ? key : (int)(0x80000000u | (unsigned)(-key));
  float out;
#ifdef __CUDA_ARCH__
  out = __int_as_float(ob);
#else
  memcpy(&out, &ob, 4);
#endif
  return out;
}

// _misc.py:71-82
template <class R> __device__ __forceinline__ R linear_rescale(R t0, R t, R t1) {
  const bool cond = (t0 == t1);
  const R num = cond ? R(0) : t - t0;
  const R den = cond ? R(1) : t1 - t0;
  return num / den;
}

// ---- fast, division-free building blocks for the PID controller's hot path ----
// The step-size factor does not need to be correctly rounded: a relative error eps in dt changes the state by
// ~order * eps * (local error) (1e-13 relative here moves a 1e-8-accurate step by ~1e-20), and the accept/reject
// decision is protected separately (the kernel sends anything within 1e-6 of the boundary to the faithful path).
// So one Newton step on each SFU seed (~1e-13) is ample; CUDA's pow() / IEEE division would cost ~10x more.
// 1/d for positive normal d: MUFU.RCP64H seed (rel. err <= 2^-23) + one Newton step -> ~2^-46.
__device__ __forceinline__ double fast_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  const double e = fma(-d, r, 1.0);
  return fma(r, e, r);
}
__device__ __forceinline__ float fast_rcp(float d) { return __frcp_rn(d); }

template <int M> __device__ __forceinline__ double ipow(double z) {
  if constexpr (M == 1) return z;
  else if constexpr (M % 2 == 0) { const double h = ipow<M / 2>(z); return h * h; }
  else return ipow<M - 1>(z) * z;
}
template <int M> __device__ __forceinline__ float ipowf(float z) {
  if constexpr (M == 1) return z;
  else if constexpr (M % 2 == 0) { const float h = ipowf<M / 2>(z); return h * h; }
  else return ipowf<M - 1>(z) * z;
}
// q^(-1/M) for q in [1e-30, 1e30] (qf == (float)q): fp32 SFU seed (MUFU.LG2 / MUFU.EX2, ~1e-6) refined by one
// division-free Newton step z <- z + z (1 - q z^M) / M in fp32 (FMA pipe, -> ~1e-7) and one in fp64
// (error e -> (M+1)/2 e^2 ~ 1e-13): log2(M) + 4 FP64 instructions.
template <int M> __device__ __forceinline__ double inv_root(double q, float qf) {
  float s;
  float lg;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(qf));  // qf is a normal number: no denormal fix-up needed
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(lg * (-1.0f / (float)M)));
  s = fmaf(s * fmaf(-qf, ipowf<M>(s), 1.0f), 1.0f / (float)M, s);
  double z = (double)s;
  const double r = fma(-q, ipow<M>(z), 1.0);
  return fma(z * r, 1.0 / (double)M, z);
}

// max(|a|, |b|) and min / max of POSITIVE doubles on the integer ALU (the IEEE bit patterns of non-negative
// doubles order like the values; NaN has the largest pattern and therefore propagates through max).
// Keeps these off the FP64 pipe, where DSETP + select + NaN fix-up would cost an FP64 issue slot each.
__device__ __forceinline__ double abs_max_bits(double a, double b) {
  // 32-bit halves on purpose: masking the 64-bit pattern is turned back into DADD |x| (an FP64-pipe op) by ptxas
  const int ah = __double2hiint(a) & 0x7fffffff, bh = __double2hiint(b) & 0x7fffffff;
  const unsigned al = (unsigned)__double2loint(a), bl = (unsigned)__double2loint(b);
  const bool gt = (ah > bh) || (ah == bh && al > bl);
  return __hiloint2double(gt ? ah : bh, (int)(gt ? al : bl));
}
__device__ __forceinline__ double pos_max_bits(double a, double b) {
  const long long x = __double_as_longlong(a), y = __double_as_longlong(b);
  return __longlong_as_double(x > y ? x : y);
}
__device__ __forceinline__ double pos_min_bits(double a, double b) {
  const long long x = __double_as_longlong(a), y = __double_as_longlong(b);
  return __longlong_as_double(x < y ? x : y);
}
__device__ __forceinline__ float abs_max_bits(float a, float b) { return fmaxf(fabsf(a), fabsf(b)); }
__device__ __forceinline__ float pos_max_bits(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ float pos_min_bits(float a, float b) { return fminf(a, b); }

// streaming (evict-first) stores for write-once outputs
__device__ __forceinline__ void st_cs(double *p, double v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs(float *p, float v) { __stcs(p, v); }
// 32 bytes (STG.E.256, new on sm_100) of one value / of four or eight given values; p must be 32-byte aligned
__device__ __forceinline__ void st32B_fill_cs(double *p, double v) {
  asm volatile("st.global.cs.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void st32B_fill_cs(float *p, float v) {
  asm volatile("st.global.cs.v8.f32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st32B(double *p, const double *v, bool cs) {  // v: 4 values (shared memory, 16-byte aligned)
  const double2 a = *reinterpret_cast<const double2 *>(v), b = *reinterpret_cast<const double2 *>(v + 2);
  if (cs) asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
  else asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
}
__device__ __forceinline__ void st32B(float *p, const float *v, bool cs) {    // v: 8 values
  const float4 a = *reinterpret_cast<const float4 *>(v), b = *reinterpret_cast<const float4 *>(v + 4);
  if (cs) asm volatile("st.global.cs.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
  else asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}

// Store a register-resident row of N values with the widest vector store its size allows: 256-bit
// (STG.E.256, new on sm_100), 128-bit, else scalar.  `vec_ok` says the destination rows are 32-byte aligned
// (checked on the host from the base pointer and the row stride); every store is a streaming (.cs) store.
// A lane writing a contiguous row with 256-bit stores covers whole 32-byte sectors by itself, so the L2 sees
// full-sector writes even though neighbouring lanes write to different trajectories' rows.
template <int N>
__device__ __forceinline__ void store_row(double *dst, const double (&v)[N], bool vec_ok, bool cs = true) {
  if (!cs) {  // write-back stores (dense records: L2 merges neighbouring steps)
    if (vec_ok && N % 4 == 0) {
#pragma unroll
      for (int i = 0; i < N; i += 4)
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "d"(v[i]), "d"(v[i + 1 < N ? i + 1 : i]),
                     "d"(v[i + 2 < N ? i + 2 : i]), "d"(v[i + 3 < N ? i + 3 : i]) : "memory");
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) dst[i] = v[i];
    }
    return;
  }
  if (vec_ok && N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N; i += 4)
      asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "d"(v[i]), "d"(v[i + 1 < N ? i + 1 : i]),
                   "d"(v[i + 2 < N ? i + 2 : i]), "d"(v[i + 3 < N ? i + 3 : i]) : "memory");
  } else if (vec_ok && N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N; i += 2)
      asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(dst + i), "d"(v[i]), "d"(v[i + 1 < N ? i + 1 : i]) : "memory");
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) __stcs(dst + i, v[i]);
  }
}
template <int N>
__device__ __forceinline__ void store_row(float *dst, const float (&v)[N], bool vec_ok, bool cs = true) {
  if (!cs) {
#pragma unroll
    for (int i = 0; i < N; ++i) dst[i] = v[i];
    return;
  }
  if (vec_ok && N % 8 == 0) {
#pragma unroll
    for (int i = 0; i < N; i += 8)
      asm volatile("st.global.cs.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + i), "f"(v[i]),
                   "f"(v[i + 1 < N ? i + 1 : i]), "f"(v[i + 2 < N ? i + 2 : i]), "f"(v[i + 3 < N ? i + 3 : i]),
                   "f"(v[i + 4 < N ? i + 4 : i]), "f"(v[i + 5 < N ? i + 5 : i]), "f"(v[i + 6 < N ? i + 6 : i]),
                   "f"(v[i + 7 < N ? i + 7 : i]) : "memory");
  } else if (vec_ok && N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N; i += 4)
      asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(v[i]), "f"(v[i + 1 < N ? i + 1 : i]),
                   "f"(v[i + 2 < N ? i + 2 : i]), "f"(v[i + 3 < N ? i + 3 : i]) : "memory");
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) __stcs(dst + i, v[i]);
  }
}

}  // namespace dfx
