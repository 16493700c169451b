// vbt.cuh - VirtualBrownianTree on the device (one leaf of shape () or (m,)).
//
// Replaces diffrax/_brownian/tree.py: _evaluate_leaf (366-624: root draw 377-403, descent loop
// 409-448, final "sqrt"-spline bridge 563-621), _brownian_arch (626-773), _levy_diff (94-148)
// and _denormalise_bm_inc (303-323).  One thread walks one tree; the depth
// L = ceil(log2(1/tol_normalised)) is the same for every trajectory (it depends only on tol),
// so the descent loop is warp-uniform and only the left/right selects differ between lanes.
//
// Per level only the blocks that are needed are generated: the midpoint key, the chosen child
// key (going RIGHT selects key_st, going LEFT key_tu - tree.py:431-432) and the normals.
#pragma once
#include "prng.cuh"

namespace dfx {

template <class R, int M> struct LevyVal { R dt, W[M], H[M], barH[M]; };

struct VbtParams {
  double t0, t1;   // tree interval (ctor arguments)
  int depth;       // number of descent levels: smallest L with 2^-L <= tol/(t1-t0)
  int levy;        // dfx_levy
  int partitionable;
  // descent cache (see VbtCache): number of tree levels whose entry state is kept per thread, and the slot stride
  // (= threads per CTA); 0 levels = no cache.  Filled by the launcher together with the dynamic shared memory size.
  int cache_levels, cache_stride;
};

// Descent cache.  _evaluate_leaf is a pure function of (key, r); two queries share every level of the descent on which
// they branch the same way, and consecutive solver steps query neighbouring times.  Per thread, shared memory keeps the
// state at the ENTRY of each level of the last descent (key, s, w_s[M], w_su[M] [, bhh_s[M], bhh_su[M]]) and the branch
// taken; a new query replays the cached branches (one load + one compare per level), resumes from the first level where
// it branches differently, and overwrites the cache from there.  The values are those of a full descent, bit for bit.
// Layout: slot (level, word) of a thread at base[(level * kWords + word) * stride], base already offset by the thread.
template <class R, bool STLA, int M> struct VbtCache {
  static constexpr int kWords = 3 + (STLA ? 4 : 2) * M;
  R *base = nullptr;
  int levels = 0, stride = 0;
  int n_entry = 0;      // entry states of levels 0 .. n_entry-1 are valid
  uint32_t path = 0;    // bit j: the cached descent went right at level j (valid for j < n_entry - 1)
  __device__ __forceinline__ R &slot(int level, int word) const { return base[(level * kWords + word) * stride]; }
};
template <class R> __device__ __forceinline__ R word_from_u32(uint32_t u);
template <> __device__ __forceinline__ float word_from_u32<float>(uint32_t u) { return __uint_as_float(u); }
template <> __device__ __forceinline__ double word_from_u32<double>(uint32_t u) { return __longlong_as_double((long long)u); }
__device__ __forceinline__ uint32_t word_to_u32(float v) { return __float_as_uint(v); }
__device__ __forceinline__ uint32_t word_to_u32(double v) { return (uint32_t)__double_as_longlong(v); }

// 2^-level as R, exact
template <class R> __device__ __forceinline__ R pow2_neg(int level);
template <> __device__ __forceinline__ double pow2_neg<double>(int level) {
  return __longlong_as_double((long long)(1023 - level) << 52);
}
template <> __device__ __forceinline__ float pow2_neg<float>(int level) { return __int_as_float((127 - level) << 23); }

template <class R> __device__ __forceinline__ R relu(R x) { return (x != x) ? x : (x > R(0) ? x : R(0)); }

// tree.py:366-624 with `leaf_key` = split_by_tree(user_key, shape)[0] (tree.py:301).  M = number of components of the
// leaf: shape () -> 1, shape (m,) -> m (ONE leaf either way: the components share the key path, each node draws
// jr.normal(key, shape)).  Every float operation is an explicitly rounded x_* (prng.cuh): the oracle performs the same
// sequence, so the values agree bit for bit.
template <class R, bool STLA, int M>
__device__ __forceinline__ LevyVal<R, M> vbt_evaluate_leaf(Key leaf_key, R r, const VbtParams &vp, VbtCache<R, STLA, M> *cache = nullptr) {
  const bool part = vp.partitionable != 0;
  Key key;
  R w_s[M], w_su[M], bhh_s[M], bhh_su[M];
#pragma unroll
  for (int c = 0; c < M; ++c) { w_s[c] = R(0); w_su[c] = R(0); bhh_s[c] = R(0); bhh_su[c] = R(0); }
  R s = R(0);
  int level0 = 0;
  const bool use_cache = cache != nullptr && cache->levels > 0;
  auto cache_store = [&](int e) {
    cache->slot(e, 0) = word_from_u32<R>(key.a);
    cache->slot(e, 1) = word_from_u32<R>(key.b);
    cache->slot(e, 2) = s;
#pragma unroll
    for (int c = 0; c < M; ++c) {
      cache->slot(e, 3 + c) = w_s[c]; cache->slot(e, 3 + M + c) = w_su[c];
      if constexpr (STLA) { cache->slot(e, 3 + 2 * M + c) = bhh_s[c]; cache->slot(e, 3 + 3 * M + c) = bhh_su[c]; }
    }
  };
  if (use_cache && cache->n_entry > 0) {
    // replay the cached branches while this query takes the same ones
    int j = 0;
    while (j < cache->n_entry - 1) {
      const R sj = cache->slot(j, 2);
      const bool right = r > sj + pow2_neg<R>(j + 1);
      if (right != (((cache->path >> j) & 1u) != 0u)) break;
      ++j;
    }
    level0 = j;
    key.a = word_to_u32(cache->slot(j, 0));
    key.b = word_to_u32(cache->slot(j, 1));
    s = cache->slot(j, 2);
#pragma unroll
    for (int c = 0; c < M; ++c) {
      w_s[c] = cache->slot(j, 3 + c); w_su[c] = cache->slot(j, 3 + M + c);
      if constexpr (STLA) { bhh_s[c] = cache->slot(j, 3 + 2 * M + c); bhh_su[c] = cache->slot(j, 3 + 3 * M + c); }
    }
    cache->n_entry = j + 1;
    cache->path &= (j >= 32) ? 0xFFFFFFFFu : ((1u << j) - 1u);
  } else {
    if constexpr (STLA) {  // state_key, init_key_w, init_key_hh = split(key, 3)
      key = split_child<3>(leaf_key, 0, part);
      const Key kw = split_child<3>(leaf_key, 1, part), kh = split_child<3>(leaf_key, 2, part);
#pragma unroll
      for (int c = 0; c < M; ++c) {
        w_su[c] = random_normal<R, M>(kw, part, c);
        bhh_su[c] = x_div(random_normal<R, M>(kh, part, c), R(3.4641016151377544));  // math.sqrt(12)
      }
    } else {               // state_key, init_key_w = split(key, 2)
      key = split_child<2>(leaf_key, 0, part);
      const Key kw = split_child<2>(leaf_key, 1, part);
#pragma unroll
      for (int c = 0; c < M; ++c) w_su[c] = random_normal<R, M>(kw, part, c);
    }
    if (use_cache) {
      cache_store(0);
      cache->n_entry = 1;
      cache->path = 0;
    }
  }
  for (int level = level0; level < vp.depth; ++level) {
    const R su = pow2_neg<R>(level);
    const R st = pow2_neg<R>(level + 1);  // su / 2
    const R t = x_add(s, st);
    const R root_su = x_sqrt(su);
    [[maybe_unused]] const R inv_su = pow2_neg<R>(-level), su4 = pow2_neg<R>(level + 2);  // 1 / su, su / 4
    const Key mid = split_child<3>(key, 1, part);
    R w_st[M], w_tu[M], w_t[M], bhh_st[M], bhh_tu[M], bhh_t[M];
    [[maybe_unused]] Key zk1, zk2;
    if constexpr (STLA) { zk1 = split_child<2>(mid, 0, part); zk2 = split_child<2>(mid, 1, part); }
#pragma unroll
    for (int c = 0; c < M; ++c) {
      bhh_st[c] = bhh_tu[c] = bhh_t[c] = R(0);
      if constexpr (STLA) {  // tree.py:727-756
        const R z1 = random_normal<R, M>(zk1, part, c);
        const R z2 = random_normal<R, M>(zk2, part, c);
        const R z = x_mul(z1, x_mul(root_su, R(0.25)));                 // root_su / 4 (power-of-two divisors: exact either way)
        const R n = x_mul(z2, x_sqrt(x_div(su, R(12))));
        const R w_term1 = x_mul(w_su[c], R(0.5));                        // w_su / 2
        const R w_term2 = x_add(x_mul(x_mul(R(1.5), inv_su), bhh_su[c]), z);  // 3 / (2 su) = 1.5 * 2^level, exact
        w_st[c] = x_add(w_term1, w_term2);
        w_tu[c] = x_sub(w_term1, w_term2);
        const R bhh_term1 = x_sub(x_mul(bhh_su[c], R(0.125)), x_mul(su4, z));  // bhh_su / 8 - su / 4 * z
        const R bhh_term2 = x_mul(su4, n);
        bhh_st[c] = x_add(bhh_term1, bhh_term2);
        bhh_tu[c] = x_sub(bhh_term1, bhh_term2);
        w_t[c] = x_add(w_s[c], w_st[c]);
        bhh_t[c] = x_add(x_add(bhh_s[c], bhh_st[c]), x_mul(R(0.5), x_sub(x_mul(t, w_s[c]), x_mul(s, w_t[c]))));
      } else {               // tree.py:758-768
        const R mean = x_mul(R(0.5), w_su[c]);
        const R w_term2 = x_mul(x_mul(root_su, R(0.5)), random_normal<R, M>(mid, part, c));
        w_st[c] = x_add(mean, w_term2);
        w_tu[c] = x_sub(mean, w_term2);
        w_t[c] = x_add(w_s[c], w_st[c]);
      }
    }
    const bool right = r > t;  // tree.py:429-437, _split_interval 162-174
    key = split_child<3>(key, right ? 0 : 2, part);
    s = right ? t : s;
#pragma unroll
    for (int c = 0; c < M; ++c) {
      w_s[c] = right ? w_t[c] : w_s[c];
      w_su[c] = right ? w_tu[c] : w_st[c];
      if constexpr (STLA) {
        bhh_s[c] = right ? bhh_t[c] : bhh_s[c];
        bhh_su[c] = right ? bhh_tu[c] : bhh_st[c];
      }
    }
    if (use_cache && level + 1 < cache->levels && level < 32) {  // entry state of level + 1, and the branch just taken
      cache_store(level + 1);
      cache->n_entry = level + 2;
      cache->path |= right ? (1u << level) : 0u;
    }
  }
  // tree.py:450-455
  const R su = pow2_neg<R>(vp.depth);
  const R inv_su = pow2_neg<R>(-vp.depth);  // x / su == x * inv_su exactly (power of two)
  const R sr = relu(x_sub(r, s));
  const R ru = relu(x_sub(su, sr));
  LevyVal<R, M> out;
  out.dt = r;
  if constexpr (STLA) {  // tree.py:563-605
    const R sr3 = x_mul(x_mul(sr, sr), sr), ru3 = x_mul(x_mul(ru, ru), ru);
    const R inv_su3 = pow2_neg<R>(-3 * vp.depth);
    const Key k1 = split_child<2>(key, 0, part), k2 = split_child<2>(key, 1, part);
    const R sr_ru_half = x_sqrt(x_mul(sr, ru));
    const R d = x_sqrt(x_add(sr3, ru3));
    const R d_prime = x_div(R(1), x_mul(x_mul(R(2), su), d));
    const R a = x_mul(x_mul(d_prime, sr3), sr_ru_half);
    const R b = x_mul(x_mul(d_prime, ru3), sr_ru_half);
    const R cc = x_div(x_sqrt(x_mul(x_mul(R(3), sr3), ru3)), x_mul(R(6), d));
    const R inverse_r = x_div(R(1), r_abs(r) < Num<R>::eps() ? Num<R>::inf() : r);
    const R c_w = x_mul(sr, inv_su);                                   // sr / su
    const R c_h = x_mul(x_mul(x_mul(R(6), sr), ru), inv_su3);          // 6 sr ru / su^3
    const R c_x = x_mul(x_mul(R(2), x_add(a, b)), inv_su);             // 2 (a + b) / su
    const R c_b = x_mul(sr3, inv_su3);                                 // sr^3 / su^3
#pragma unroll
    for (int c = 0; c < M; ++c) {
      const R x1 = random_normal<R, M>(k1, part, c);
      const R x2 = random_normal<R, M>(k2, part, c);
      const R w_sr = x_add(x_add(x_mul(c_w, w_su[c]), x_mul(c_h, bhh_su[c])), x_mul(c_x, x1));
      const R w_r = x_add(w_s[c], w_sr);
      const R bhh_sr = x_add(x_sub(x_mul(c_b, bhh_su[c]), x_mul(a, x1)), x_mul(cc, x2));
      const R bhh_r = x_add(x_add(bhh_s[c], bhh_sr), x_mul(R(0.5), x_sub(x_mul(r, w_s[c]), x_mul(s, w_r))));
      out.W[c] = w_r;
      out.barH[c] = bhh_r;
      out.H[c] = x_mul(inverse_r, bhh_r);
    }
  } else {               // tree.py:607-621
    const R scale = x_sqrt(x_mul(x_mul(sr, ru), inv_su));  // sqrt(sr * ru / su)
    const R frac = x_mul(sr, inv_su);                      // sr / su
#pragma unroll
    for (int c = 0; c < M; ++c) {
      const R w_mean = x_add(w_s[c], x_mul(frac, w_su[c]));
      const R z = random_normal<R, M>(key, part, c);
      out.W[c] = x_add(w_mean, x_mul(scale, z));
      out.H[c] = R(0);
      out.barH[c] = R(0);
    }
  }
  return out;
}

// A per-trajectory tree with a two-entry memo: consecutive steps share an endpoint
// (this step's t0 is the previous step's t1, or - after a rejection - its t0), and
// _evaluate_leaf is a pure function of (key, r), so reusing the stored value is bit-identical
// to re-walking the tree.
template <class R, bool STLA, int M = 1>
struct BrownianTree {
  Key leaf;
  R T0, T1, sqrt_len;
  R memo_t[2];
  LevyVal<R, M> memo_v[2];
  VbtCache<R, STLA, M> cache;

  // the descent cache lives in dynamic shared memory: `smem` = the CTA's cache area, this thread uses column threadIdx.x
  __device__ __forceinline__ void attach_cache(R *smem, const VbtParams &vp) {
    cache.base = smem + threadIdx.x;
    cache.levels = vp.cache_levels;
    cache.stride = vp.cache_stride;
  }

  __device__ __forceinline__ void init(const uint32_t *user_key, const VbtParams &vp) {
    cache.n_entry = 0;
    cache.path = 0;
    Key k{user_key[0], user_key[1]};
    leaf = split_child<1>(k, 0, vp.partitionable != 0);  // split_by_tree(key, shape) == split(key, 1)[0]: one leaf for () and (m,)
    T0 = (R)vp.t0;
    T1 = (R)vp.t1;
    sqrt_len = x_sqrt(x_sub(T1, T0));
    memo_t[0] = memo_t[1] = Num<R>::nan();
  }

  __device__ __forceinline__ LevyVal<R, M> at(R t, const VbtParams &vp) {
    if (t == memo_t[1]) return memo_v[1];
    if (t == memo_t[0]) return memo_v[0];
    return vbt_evaluate_leaf<R, STLA, M>(leaf, linear_rescale(T0, t, T1), vp, &cache);
  }

  // evaluate(ta, tb, use_levy=True): tree.py:326-354 + _levy_diff + _denormalise_bm_inc
  __device__ __forceinline__ void increment(R ta, R tb, const VbtParams &vp, R (&W)[M], R (&H)[M]) {
    const LevyVal<R, M> x0 = at(ta, vp);
    const LevyVal<R, M> x1 = at(tb, vp);
    memo_t[0] = ta; memo_v[0] = x0;
    memo_t[1] = tb; memo_v[1] = x1;
    const R su = x_sub(x1.dt, x0.dt);
    [[maybe_unused]] const R inverse_su = x_div(R(1), r_abs(su) < Num<R>::eps() ? Num<R>::inf() : su);
#pragma unroll
    for (int c = 0; c < M; ++c) {
      const R w_su = x_sub(x1.W[c], x0.W[c]);
      W[c] = x_mul(sqrt_len, w_su);
      if constexpr (STLA) {
        const R u_bb_s = x_sub(x_mul(x1.dt, x0.W[c]), x_mul(x0.dt, x1.W[c]));
        const R bhh_su = x_sub(x_sub(x1.barH[c], x0.barH[c]), x_mul(R(0.5), u_bb_s));
        H[c] = x_mul(sqrt_len, x_mul(inverse_su, bhh_su));
      } else {
        H[c] = R(0);
      }
    }
  }
};

}  // namespace dfx
