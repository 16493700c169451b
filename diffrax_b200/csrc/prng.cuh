// prng.cuh - JAX's threefry2x32 PRNG on the device, bit-exact on the integer side.
//
// Replaces jax.random.{split, normal} as used by diffrax/_brownian/tree.py:378-402, 575-577,
// 611, 666, 732-734, 762 and _misc.py:133.  JAX itself is not under /root/reference; the
// block function is the published Threefry-2x32-20 (Salmon et al., SC'11) and is pinned by
// the Random123 known-answer vectors; the key/counter layouts follow
// jax/_src/prng.py for both values of `jax_threefry_partitionable`.
#pragma once
#include "common.cuh"

namespace dfx {

struct Key { uint32_t a, b; };

__device__ __forceinline__ uint32_t rotl(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

// One Threefry-2x32 block, 20 rounds: 5 groups of 4 (add, rotate, xor) with a key injection
// after each group.  Pure INT32 ALU work: 20 IADD + 20 SHF + 20 LOP3 + ~12 injection adds.
__device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1,
                                             uint32_t &o0, uint32_t &o1) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
  x0 += k0; x1 += k1;
#define DFX_TF_ROUND(r) x0 += x1; x1 = rotl(x1, r); x1 ^= x0;
  DFX_TF_ROUND(13) DFX_TF_ROUND(15) DFX_TF_ROUND(26) DFX_TF_ROUND(6)
  x0 += k1; x1 += k2 + 1u;
  DFX_TF_ROUND(17) DFX_TF_ROUND(29) DFX_TF_ROUND(16) DFX_TF_ROUND(24)
  x0 += k2; x1 += k0 + 2u;
  DFX_TF_ROUND(13) DFX_TF_ROUND(15) DFX_TF_ROUND(26) DFX_TF_ROUND(6)
  x0 += k0; x1 += k1 + 3u;
  DFX_TF_ROUND(17) DFX_TF_ROUND(29) DFX_TF_ROUND(16) DFX_TF_ROUND(24)
  x0 += k1; x1 += k2 + 4u;
  DFX_TF_ROUND(13) DFX_TF_ROUND(15) DFX_TF_ROUND(26) DFX_TF_ROUND(6)
  x0 += k2; x1 += k0 + 5u;
#undef DFX_TF_ROUND
  o0 = x0; o1 = x1;
}

// jax.random.split(key, NUM)[i] for compile-time NUM and i, computing only the blocks that
// feed child i.
//   partitionable: child i = block(0, i)                                    -> 1 block
//   original:      flat = concat(y0[0..NUM), y1[0..NUM)) with (y0[j], y1[j]) = block(j, NUM+j);
//                  child i = (flat[2i], flat[2i+1])                         -> 1 or 2 blocks
template <int NUM>
__device__ __forceinline__ Key split_child(Key key, int i, bool partitionable) {
  Key out;
  if (partitionable) {
    threefry2x32(key.a, key.b, 0u, (uint32_t)i, out.a, out.b);
    return out;
  }
  uint32_t w[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int f = 2 * i + h;           // flat index
    const int j = f < NUM ? f : f - NUM;
    uint32_t y0, y1;
    threefry2x32(key.a, key.b, (uint32_t)j, (uint32_t)(NUM + j), y0, y1);
    w[h] = f < NUM ? y0 : y1;
  }
  out.a = w[0]; out.b = w[1];
  return out;
}

// ---- explicitly rounded arithmetic ----
// nvcc contracts a*b+c into FMA on its own; the Brownian path must not depend on that choice.  These spell every
// rounding out (the _rn intrinsics are never contracted), so the float side of the PRNG / VirtualBrownianTree is ONE
// defined sequence of IEEE-754 operations - the same sequence the CPU oracle performs (oracle/oracle.c, built with
// -ffp-contract=off): Brownian increments agree bit for bit between the two.
__device__ __forceinline__ double x_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double x_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double x_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double x_div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double x_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ double x_sqrt(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ float x_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float x_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float x_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float x_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float x_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float x_sqrt(float a) { return __fsqrt_rn(a); }

// random bits of element `w` of a draw of shape (M,) (prng.py threefry_random_bits; M = 1 is shape ()):
//   partitionable: counter (0, w): 32-bit -> x0 ^ x1; 64-bit -> (x0 << 32) | x1
//   original 32-bit: counters iota(M) padded with one 0 to even length 2h, halves (c[:h], c[h:]); words are
//                    concat(x0s, x1s)[:M]: element w < h is x0 of block (w, h+w or 0 for the pad), w >= h is x1 of block (w-h, w)
//   original 64-bit: 2M words from iota(2M), halves (c[:M], c[M:]); element w is (x0 << 32) | x1 of block (w, M+w)
template <int M>
__device__ __forceinline__ uint32_t random_bits32(Key key, int w, bool partitionable) {
  uint32_t a, b;
  if (partitionable) { threefry2x32(key.a, key.b, 0u, (uint32_t)w, a, b); return a ^ b; }
  constexpr int h = (M + 1) / 2;
  if (w < h) { threefry2x32(key.a, key.b, (uint32_t)w, (h + w < M) ? (uint32_t)(h + w) : 0u, a, b); return a; }
  threefry2x32(key.a, key.b, (uint32_t)(w - h), (uint32_t)w, a, b);
  return b;
}
template <int M>
__device__ __forceinline__ unsigned long long random_bits64(Key key, int w, bool partitionable) {
  uint32_t a, b;
  threefry2x32(key.a, key.b, partitionable ? 0u : (uint32_t)w, partitionable ? (uint32_t)w : (uint32_t)(M + w), a, b);
  return ((unsigned long long)a << 32) | (unsigned long long)b;
}

// log1p on [-1, 0] (x = -u*u is all erf_inv needs) as ONE explicit, branch-free operation sequence - the same one as
// oracle/oracle.c orc_log1p_*: 1+x = 2^k (1+f), f in [sqrt(1/2)-1, sqrt(2)-1), taken from the exact x by one fma
// (f = 2^-k x + (2^-k - 1)); log(1+f) = f - f^2/2 + s (f^2/2 + R(s^2)), s = f/(2+f) (the published fdlibm scheme and
// coefficients).  <= 1 ulp.
__device__ __forceinline__ double x_log1p(double x) {
  constexpr double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
  if (!(x > -1.0 && x <= 0.0)) return x == -1.0 ? -Num<double>::inf() : Num<double>::nan();
  const double u = x_add(1.0, x);
  const int hi = __double2hiint(u);
  const unsigned lo = (unsigned)__double2loint(u);
  const int mh = hi & 0x000fffff;  // mantissa >= that of sqrt(2) (0x6a09e 667f3bcd)?
  int k = ((hi >> 20) & 0x7ff) - 1023 + ((mh > 0x6a09e || (mh == 0x6a09e && lo >= 0x667f3bcdu)) ? 1 : 0);
  k = (x > -0.2928932188134524) ? 0 : k;
  const double scale = __hiloint2double((1023 - k) << 20, 0);
  const double f = x_fma(scale, x, x_sub(scale, 1.0));
  const double hfsq = x_mul(x_mul(0.5, f), f);
  const double s = x_div(f, x_add(2.0, f));
  const double z = x_mul(s, s);
  double r = 1.479819860511658591e-01;
  r = x_fma(r, z, 1.531383769920937332e-01);
  r = x_fma(r, z, 1.818357216161805012e-01);
  r = x_fma(r, z, 2.222219843214978396e-01);
  r = x_fma(r, z, 2.857142874366239149e-01);
  r = x_fma(r, z, 3.999999999940941908e-01);
  r = x_fma(r, z, 6.666666666666735130e-01);
  r = x_mul(r, z);
  const double dk = (double)k;
  return x_sub(x_mul(dk, ln2_hi), x_sub(x_sub(hfsq, x_add(x_mul(s, x_add(hfsq, r)), x_mul(dk, ln2_lo))), f));
}
__device__ __forceinline__ float x_log1p(float x) {
  constexpr float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f;
  if (!(x > -1.0f && x <= 0.0f)) return x == -1.0f ? -Num<float>::inf() : Num<float>::nan();
  const float u = x_add(1.0f, x);
  const int b = __float_as_int(u);
  int k = ((b >> 23) & 0xff) - 127 + (((b & 0x007fffff) >= 0x3504f3) ? 1 : 0);
  k = (x > -0.29289323f) ? 0 : k;
  const float scale = __int_as_float((127 - k) << 23);
  const float f = x_fma(scale, x, x_sub(scale, 1.0f));
  const float hfsq = x_mul(x_mul(0.5f, f), f);
  const float s = x_div(f, x_add(2.0f, f));
  const float z = x_mul(s, s);
  float r = 1.4798198640e-01f;
  r = x_fma(r, z, 1.5313838422e-01f);
  r = x_fma(r, z, 1.8183572590e-01f);
  r = x_fma(r, z, 2.2222198546e-01f);
  r = x_fma(r, z, 2.8571429849e-01f);
  r = x_fma(r, z, 4.0000000596e-01f);
  r = x_fma(r, z, 6.6666668653e-01f);
  r = x_mul(r, z);
  const float dk = (float)k;
  return x_sub(x_mul(dk, ln2_hi), x_sub(x_sub(hfsq, x_add(x_mul(s, x_add(hfsq, r)), x_mul(dk, ln2_lo))), f));
}

// lax.erf_inv - Giles' polynomials in w = -log1p(-x*x), coefficients and branches as XLA evaluates them
// (ErfInv32/64).  [EXT: restated from the published algorithm.]  Sequence: one fma per Horner step, result p * x.
__device__ __forceinline__ float erfinv_xla(float x) {
  float w = -x_log1p(-x_mul(x, x));
  float p;
  if (w < 5.0f) {
    w = x_sub(w, 2.5f);
    p = 2.81022636e-08f;
    p = x_fma(p, w, 3.43273939e-07f);
    p = x_fma(p, w, -3.5233877e-06f);
    p = x_fma(p, w, -4.39150654e-06f);
    p = x_fma(p, w, 0.00021858087f);
    p = x_fma(p, w, -0.00125372503f);
    p = x_fma(p, w, -0.00417768164f);
    p = x_fma(p, w, 0.246640727f);
    p = x_fma(p, w, 1.50140941f);
  } else {
    w = x_sub(x_sqrt(w), 3.0f);
    p = -0.000200214257f;
    p = x_fma(p, w, 0.000100950558f);
    p = x_fma(p, w, 0.00134934322f);
    p = x_fma(p, w, -0.00367342844f);
    p = x_fma(p, w, 0.00573950773f);
    p = x_fma(p, w, -0.0076224613f);
    p = x_fma(p, w, 0.00943887047f);
    p = x_fma(p, w, 1.00167406f);
    p = x_fma(p, w, 2.83297682f);
  }
  return fabsf(x) == 1.0f ? x * Num<float>::inf() : x_mul(p, x);
}

__device__ __forceinline__ double erfinv_xla(double x) {
  double w = -x_log1p(-x_mul(x, x));
  double p;
  if (w < 6.25) {
    w = x_sub(w, 3.125);
    p = -3.6444120640178196996e-21;
    p = x_fma(p, w, -1.685059138182016589e-19);
    p = x_fma(p, w, 1.2858480715256400167e-18);
    p = x_fma(p, w, 1.115787767802518096e-17);
    p = x_fma(p, w, -1.333171662854620906e-16);
    p = x_fma(p, w, 2.0972767875968561637e-17);
    p = x_fma(p, w, 6.6376381343583238325e-15);
    p = x_fma(p, w, -4.0545662729752068639e-14);
    p = x_fma(p, w, -8.1519341976054721522e-14);
    p = x_fma(p, w, 2.6335093153082322977e-12);
    p = x_fma(p, w, -1.2975133253453532498e-11);
    p = x_fma(p, w, -5.4154120542946279317e-11);
    p = x_fma(p, w, 1.051212273321532285e-09);
    p = x_fma(p, w, -4.1126339803469836976e-09);
    p = x_fma(p, w, -2.9070369957882005086e-08);
    p = x_fma(p, w, 4.2347877827932403518e-07);
    p = x_fma(p, w, -1.3654692000834678645e-06);
    p = x_fma(p, w, -1.3882523362786468719e-05);
    p = x_fma(p, w, 0.0001867342080340571352);
    p = x_fma(p, w, -0.00074070253416626697512);
    p = x_fma(p, w, -0.0060336708714301490533);
    p = x_fma(p, w, 0.24015818242558961693);
    p = x_fma(p, w, 1.6536545626831027356);
  } else if (w < 16.0) {
    w = x_sub(x_sqrt(w), 3.25);
    p = 2.2137376921775787049e-09;
    p = x_fma(p, w, 9.0756561938885390979e-08);
    p = x_fma(p, w, -2.7517406297064545428e-07);
    p = x_fma(p, w, 1.8239629214389227755e-08);
    p = x_fma(p, w, 1.5027403968909827627e-06);
    p = x_fma(p, w, -4.013867526981545969e-06);
    p = x_fma(p, w, 2.9234449089955446044e-06);
    p = x_fma(p, w, 1.2475304481671778723e-05);
    p = x_fma(p, w, -4.7318229009055733981e-05);
    p = x_fma(p, w, 6.8284851459573175448e-05);
    p = x_fma(p, w, 2.4031110387097893999e-05);
    p = x_fma(p, w, -0.0003550375203628474796);
    p = x_fma(p, w, 0.00095328937973738049703);
    p = x_fma(p, w, -0.0016882755560235047313);
    p = x_fma(p, w, 0.0024914420961078508066);
    p = x_fma(p, w, -0.0037512085075692412107);
    p = x_fma(p, w, 0.005370914553590063617);
    p = x_fma(p, w, 1.0052589676941592334);
    p = x_fma(p, w, 3.0838856104922207635);
  } else {
    w = x_sub(x_sqrt(w), 5.0);
    p = -2.7109920616438573243e-11;
    p = x_fma(p, w, -2.5556418169965252055e-10);
    p = x_fma(p, w, 1.5076572693500548083e-09);
    p = x_fma(p, w, -3.7894654401267369937e-09);
    p = x_fma(p, w, 7.6157012080783393804e-09);
    p = x_fma(p, w, -1.4960026627149240478e-08);
    p = x_fma(p, w, 2.9147953450901080826e-08);
    p = x_fma(p, w, -6.7711997758452339498e-08);
    p = x_fma(p, w, 2.2900482228026654717e-07);
    p = x_fma(p, w, -9.9298272942317002539e-07);
    p = x_fma(p, w, 4.5260625972231537039e-06);
    p = x_fma(p, w, -1.9681778105531670567e-05);
    p = x_fma(p, w, 7.5995277030017761139e-05);
    p = x_fma(p, w, -0.00021503011930044477347);
    p = x_fma(p, w, -0.00013871931833623122026);
    p = x_fma(p, w, 1.0103004648645343977);
    p = x_fma(p, w, 4.8499064014085844221);
  }
  return fabs(x) == 1.0 ? x * Num<double>::inf() : x_mul(p, x);
}

// jax.random.normal(key, (M,), dtype)[w] (M = 1, w = 0: shape ()): mantissa fill -> [1,2) - 1 -> u in [lo, 1) -> sqrt(2) erf_inv(u)
template <class R> __device__ __forceinline__ R normal_from_bits(typename Num<R>::uint_t bits);
template <> __device__ __forceinline__ float normal_from_bits<float>(uint32_t bits) {
  const float f = x_sub(__uint_as_float((bits >> 9) | 0x3F800000u), 1.0f);
  const float lo = -0.99999994f;               // nextafter(-1, 0)
  float u = x_fma(f, 2.0f, lo);                // f * (hi - lo) + lo, (hi - lo) rounds to 2.0f; one fma (as the oracle)
  u = fmaxf(lo, u);
  return x_mul(1.41421354f, erfinv_xla(u));    // np.array(np.sqrt(2), float32)
}
template <> __device__ __forceinline__ double normal_from_bits<double>(unsigned long long bits) {
  const double f = x_sub(__longlong_as_double((long long)((bits >> 12) | 0x3FF0000000000000ULL)), 1.0);
  const double lo = -0.99999999999999989;      // nextafter(-1, 0)
  double u = x_fma(f, 2.0, lo);                // (hi - lo) = 2 - 2^-53 rounds to 2.0
  u = fmax(lo, u);
  return x_mul(1.4142135623730951, erfinv_xla(u));
}
template <class R, int M = 1> __device__ __forceinline__ R random_normal(Key key, bool partitionable, int w = 0) {
  if constexpr (sizeof(R) == 8) return normal_from_bits<R>(random_bits64<M>(key, w, partitionable));
  else return normal_from_bits<R>(random_bits32<M>(key, w, partitionable));
}

}  // namespace dfx
