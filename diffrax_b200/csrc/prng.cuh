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

// random_bits(key, nbits, shape=()) (prng.py threefry_random_bits)
__device__ __forceinline__ uint32_t random_bits32(Key key, bool partitionable) {
  uint32_t a, b;
  threefry2x32(key.a, key.b, 0u, 0u, a, b);
  return partitionable ? (a ^ b) : a;
}
__device__ __forceinline__ unsigned long long random_bits64(Key key, bool partitionable) {
  uint32_t a, b;
  threefry2x32(key.a, key.b, 0u, partitionable ? 0u : 1u, a, b);
  return ((unsigned long long)a << 32) | (unsigned long long)b;
}

// lax.erf_inv - Giles' polynomials in w = -log1p(-x*x), the form XLA evaluates (ErfInv32/64).
// Written with explicit fma-free Horner steps `c + p*w`; nvcc contracts them to FMA just as
// XLA's GPU backend does.  [EXT: float side is restated from the published algorithm.]
__device__ __forceinline__ float erfinv_xla(float x) {
  float w = -log1pf(-x * x);
  float p;
  if (w < 5.0f) {
    w = w - 2.5f;
    p = 2.81022636e-08f;
    p = 3.43273939e-07f + p * w;
    p = -3.5233877e-06f + p * w;
    p = -4.39150654e-06f + p * w;
    p = 0.00021858087f + p * w;
    p = -0.00125372503f + p * w;
    p = -0.00417768164f + p * w;
    p = 0.246640727f + p * w;
    p = 1.50140941f + p * w;
  } else {
    w = sqrtf(w) - 3.0f;
    p = -0.000200214257f;
    p = 0.000100950558f + p * w;
    p = 0.00134934322f + p * w;
    p = -0.00367342844f + p * w;
    p = 0.00573950773f + p * w;
    p = -0.0076224613f + p * w;
    p = 0.00943887047f + p * w;
    p = 1.00167406f + p * w;
    p = 2.83297682f + p * w;
  }
  return fabsf(x) == 1.0f ? x * Num<float>::inf() : p * x;
}

__device__ __forceinline__ double erfinv_xla(double x) {
  double w = -log1p(-x * x);
  double p;
  if (w < 6.25) {
    w = w - 3.125;
    p = -3.6444120640178196996e-21;
    p = -1.685059138182016589e-19 + p * w;
    p = 1.2858480715256400167e-18 + p * w;
    p = 1.115787767802518096e-17 + p * w;
    p = -1.333171662854620906e-16 + p * w;
    p = 2.0972767875968561637e-17 + p * w;
    p = 6.6376381343583238325e-15 + p * w;
    p = -4.0545662729752068639e-14 + p * w;
    p = -8.1519341976054721522e-14 + p * w;
    p = 2.6335093153082322977e-12 + p * w;
    p = -1.2975133253453532498e-11 + p * w;
    p = -5.4154120542946279317e-11 + p * w;
    p = 1.051212273321532285e-09 + p * w;
    p = -4.1126339803469836976e-09 + p * w;
    p = -2.9070369957882005086e-08 + p * w;
    p = 4.2347877827932403518e-07 + p * w;
    p = -1.3654692000834678645e-06 + p * w;
    p = -1.3882523362786468719e-05 + p * w;
    p = 0.0001867342080340571352 + p * w;
    p = -0.00074070253416626697512 + p * w;
    p = -0.0060336708714301490533 + p * w;
    p = 0.24015818242558961693 + p * w;
    p = 1.6536545626831027356 + p * w;
  } else if (w < 16.0) {
    w = sqrt(w) - 3.25;
    p = 2.2137376921775787049e-09;
    p = 9.0756561938885390979e-08 + p * w;
    p = -2.7517406297064545428e-07 + p * w;
    p = 1.8239629214389227755e-08 + p * w;
    p = 1.5027403968909827627e-06 + p * w;
    p = -4.013867526981545969e-06 + p * w;
    p = 2.9234449089955446044e-06 + p * w;
    p = 1.2475304481671778723e-05 + p * w;
    p = -4.7318229009055733981e-05 + p * w;
    p = 6.8284851459573175448e-05 + p * w;
    p = 2.4031110387097893999e-05 + p * w;
    p = -0.0003550375203628474796 + p * w;
    p = 0.00095328937973738049703 + p * w;
    p = -0.0016882755560235047313 + p * w;
    p = 0.0024914420961078508066 + p * w;
    p = -0.0037512085075692412107 + p * w;
    p = 0.005370914553590063617 + p * w;
    p = 1.0052589676941592334 + p * w;
    p = 3.0838856104922207635 + p * w;
  } else {
    w = sqrt(w) - 5.0;
    p = -2.7109920616438573243e-11;
    p = -2.5556418169965252055e-10 + p * w;
    p = 1.5076572693500548083e-09 + p * w;
    p = -3.7894654401267369937e-09 + p * w;
    p = 7.6157012080783393804e-09 + p * w;
    p = -1.4960026627149240478e-08 + p * w;
    p = 2.9147953450901080826e-08 + p * w;
    p = -6.7711997758452339498e-08 + p * w;
    p = 2.2900482228026654717e-07 + p * w;
    p = -9.9298272942317002539e-07 + p * w;
    p = 4.5260625972231537039e-06 + p * w;
    p = -1.9681778105531670567e-05 + p * w;
    p = 7.5995277030017761139e-05 + p * w;
    p = -0.00021503011930044477347 + p * w;
    p = -0.00013871931833623122026 + p * w;
    p = 1.0103004648645343977 + p * w;
    p = 4.8499064014085844221 + p * w;
  }
  return fabs(x) == 1.0 ? x * Num<double>::inf() : p * x;
}

// jax.random.normal(key, (), dtype): mantissa fill -> [1,2) - 1 -> u in [lo, 1) -> sqrt(2) erf_inv(u)
template <class R> __device__ __forceinline__ R random_normal(Key key, bool partitionable);
template <> __device__ __forceinline__ float random_normal<float>(Key key, bool partitionable) {
  const uint32_t bits = random_bits32(key, partitionable);
  const float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
  const float lo = -0.99999994f;               // nextafter(-1, 0)
  float u = f * (1.0f - lo) + lo;              // (hi - lo) rounds to 2.0f
  u = fmaxf(lo, u);
  return 1.41421354f * erfinv_xla(u);          // np.array(np.sqrt(2), float32)
}
template <> __device__ __forceinline__ double random_normal<double>(Key key, bool partitionable) {
  const unsigned long long bits = random_bits64(key, partitionable);
  const double f = __longlong_as_double((long long)((bits >> 12) | 0x3FF0000000000000ULL)) - 1.0;
  const double lo = -0.99999999999999989;      // nextafter(-1, 0)
  double u = f * (1.0 - lo) + lo;
  u = fmax(lo, u);
  return 1.4142135623730951 * erfinv_xla(u);
}

}  // namespace dfx
