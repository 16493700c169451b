// vbt.cuh - VirtualBrownianTree on the device (scalar leaf, shape == ()).
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

template <class R> struct LevyVal { R dt, W, H, barH; };

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
// state at the ENTRY of each level of the last descent (key, s, w_s, w_su [, bhh_s, bhh_su]) and the branch taken; a new
// query replays the cached branches (one load + one compare per level), resumes from the first level where it branches
// differently, and overwrites the cache from there.  The values are those of a full descent, bit for bit.
// Layout: slot (level, word) of a thread at base[(level * kWords + word) * stride], base already offset by the thread.
template <class R, bool STLA> struct VbtCache {
  static constexpr int kWords = STLA ? 7 : 5;
  R *base = nullptr;
  int levels = 0, stride = 0;
  int n_entry = 0;      // entry states of levels 0 .. n_entry-1 are valid
  uint32_t path = 0;    // bit j: the cached descent went right at level j (valid for j < n_entry - 1)
  __device__ __forceinline__ R &slot(int level, int word) const { return base[(level * kWords + word) * stride]; }
  static __device__ __forceinline__ R from_u32(uint32_t u);
  static __device__ __forceinline__ uint32_t to_u32(R v);
};
template <> __device__ __forceinline__ float VbtCache<float, false>::from_u32(uint32_t u) { return __uint_as_float(u); }
template <> __device__ __forceinline__ float VbtCache<float, true>::from_u32(uint32_t u) { return __uint_as_float(u); }
template <> __device__ __forceinline__ double VbtCache<double, false>::from_u32(uint32_t u) { return __longlong_as_double((long long)u); }
template <> __device__ __forceinline__ double VbtCache<double, true>::from_u32(uint32_t u) { return __longlong_as_double((long long)u); }
template <> __device__ __forceinline__ uint32_t VbtCache<float, false>::to_u32(float v) { return __float_as_uint(v); }
template <> __device__ __forceinline__ uint32_t VbtCache<float, true>::to_u32(float v) { return __float_as_uint(v); }
template <> __device__ __forceinline__ uint32_t VbtCache<double, false>::to_u32(double v) { return (uint32_t)__double_as_longlong(v); }
template <> __device__ __forceinline__ uint32_t VbtCache<double, true>::to_u32(double v) { return (uint32_t)__double_as_longlong(v); }

// 2^-level as R, exact
template <class R> __device__ __forceinline__ R pow2_neg(int level);
template <> __device__ __forceinline__ double pow2_neg<double>(int level) {
  return __longlong_as_double((long long)(1023 - level) << 52);
}
template <> __device__ __forceinline__ float pow2_neg<float>(int level) { return __int_as_float((127 - level) << 23); }

template <class R> __device__ __forceinline__ R relu(R x) { return (x != x) ? x : (x > R(0) ? x : R(0)); }

// tree.py:366-624 with `leaf_key` = split_by_tree(user_key, shape)[0] (tree.py:301)
template <class R, bool STLA>
__device__ __forceinline__ LevyVal<R> vbt_evaluate_leaf(Key leaf_key, R r, const VbtParams &vp, VbtCache<R, STLA> *cache = nullptr) {
  const bool part = vp.partitionable != 0;
  Key key;
  R w_s = R(0), w_su, bhh_s = R(0), bhh_su = R(0);
  R s = R(0);
  int level0 = 0;
  const bool use_cache = cache != nullptr && cache->levels > 0;
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
    key.a = VbtCache<R, STLA>::to_u32(cache->slot(j, 0));
    key.b = VbtCache<R, STLA>::to_u32(cache->slot(j, 1));
    s = cache->slot(j, 2);
    w_s = cache->slot(j, 3);
    w_su = cache->slot(j, 4);
    if constexpr (STLA) { bhh_s = cache->slot(j, 5); bhh_su = cache->slot(j, 6); }
    cache->n_entry = j + 1;
    cache->path &= (j >= 32) ? 0xFFFFFFFFu : ((1u << j) - 1u);
  } else {
    if constexpr (STLA) {  // state_key, init_key_w, init_key_hh = split(key, 3)
      key = split_child<3>(leaf_key, 0, part);
      w_su = random_normal<R>(split_child<3>(leaf_key, 1, part), part);
      bhh_su = random_normal<R>(split_child<3>(leaf_key, 2, part), part) / R(3.4641016151377544);  // math.sqrt(12)
    } else {               // state_key, init_key_w = split(key, 2)
      key = split_child<2>(leaf_key, 0, part);
      w_su = random_normal<R>(split_child<2>(leaf_key, 1, part), part);
    }
    if (use_cache) {
      cache->slot(0, 0) = VbtCache<R, STLA>::from_u32(key.a);
      cache->slot(0, 1) = VbtCache<R, STLA>::from_u32(key.b);
      cache->slot(0, 2) = s; cache->slot(0, 3) = w_s; cache->slot(0, 4) = w_su;
      if constexpr (STLA) { cache->slot(0, 5) = bhh_s; cache->slot(0, 6) = bhh_su; }
      cache->n_entry = 1;
      cache->path = 0;
    }
  }
  for (int level = level0; level < vp.depth; ++level) {
    const R su = pow2_neg<R>(level);
    const R st = su / R(2);
    const R t = s + st;
    const R root_su = r_sqrt(su);
    const Key mid = split_child<3>(key, 1, part);
    R w_st, w_tu, w_t, bhh_st = R(0), bhh_tu = R(0), bhh_t = R(0);
    if constexpr (STLA) {  // tree.py:727-756
      const R z1 = random_normal<R>(split_child<2>(mid, 0, part), part);
      const R z2 = random_normal<R>(split_child<2>(mid, 1, part), part);
      const R z = z1 * (root_su / R(4));
      const R n = z2 * r_sqrt(su / R(12));
      const R w_term1 = w_su / R(2);
      const R w_term2 = (R(3) / (R(2) * su)) * bhh_su + z;
      w_st = w_term1 + w_term2;
      w_tu = w_term1 - w_term2;
      const R bhh_term1 = bhh_su / R(8) - su / R(4) * z;
      const R bhh_term2 = (su / R(4)) * n;
      bhh_st = bhh_term1 + bhh_term2;
      bhh_tu = bhh_term1 - bhh_term2;
      w_t = w_s + w_st;
      bhh_t = bhh_s + bhh_st + R(0.5) * (t * w_s - s * w_t);
    } else {               // tree.py:758-768
      const R mean = R(0.5) * w_su;
      const R w_term2 = (root_su / R(2)) * random_normal<R>(mid, part);
      w_st = mean + w_term2;
      w_tu = mean - w_term2;
      w_t = w_s + w_st;
    }
    const bool right = r > t;  // tree.py:429-437, _split_interval 162-174
    key = split_child<3>(key, right ? 0 : 2, part);
    s = right ? t : s;
    w_s = right ? w_t : w_s;
    w_su = right ? w_tu : w_st;
    if constexpr (STLA) {
      bhh_s = right ? bhh_t : bhh_s;
      bhh_su = right ? bhh_tu : bhh_st;
    }
    if (use_cache && level + 1 < cache->levels && level < 32) {  // entry state of level + 1, and the branch just taken
      const int e = level + 1;
      cache->slot(e, 0) = VbtCache<R, STLA>::from_u32(key.a);
      cache->slot(e, 1) = VbtCache<R, STLA>::from_u32(key.b);
      cache->slot(e, 2) = s; cache->slot(e, 3) = w_s; cache->slot(e, 4) = w_su;
      if constexpr (STLA) { cache->slot(e, 5) = bhh_s; cache->slot(e, 6) = bhh_su; }
      cache->n_entry = e + 1;
      cache->path |= right ? (1u << level) : 0u;
    }
  }
  // tree.py:450-455
  const R su = pow2_neg<R>(vp.depth);
  const R sr = relu(r - s);
  const R ru = relu(su - sr);
  LevyVal<R> out;
  out.dt = r;
  if constexpr (STLA) {  // tree.py:563-605
    const R sr3 = sr * sr * sr, ru3 = ru * ru * ru, su3 = su * su * su;
    const R x1 = random_normal<R>(split_child<2>(key, 0, part), part);
    const R x2 = random_normal<R>(split_child<2>(key, 1, part), part);
    const R sr_ru_half = r_sqrt(sr * ru);
    const R d = r_sqrt(sr3 + ru3);
    const R d_prime = R(1) / (R(2) * su * d);
    const R a = d_prime * sr3 * sr_ru_half;
    const R b = d_prime * ru3 * sr_ru_half;
    const R w_sr = sr / su * w_su + R(6) * sr * ru / su3 * bhh_su + R(2) * (a + b) / su * x1;
    const R w_r = w_s + w_sr;
    const R c = r_sqrt(R(3) * sr3 * ru3) / (R(6) * d);
    const R bhh_sr = sr3 / su3 * bhh_su - a * x1 + c * x2;
    const R bhh_r = bhh_s + bhh_sr + R(0.5) * (r * w_s - s * w_r);
    const R inverse_r = R(1) / (r_abs(r) < Num<R>::eps() ? Num<R>::inf() : r);
    out.W = w_r;
    out.barH = bhh_r;
    out.H = inverse_r * bhh_r;
  } else {               // tree.py:607-621
    const R w_mean = w_s + sr / su * w_su;
    const R z = random_normal<R>(key, part);
    const R bb = r_sqrt(sr * ru / su) * z;
    out.W = w_mean + bb;
    out.H = R(0);
    out.barH = R(0);
  }
  return out;
}

// A per-trajectory tree with a two-entry memo: consecutive steps share an endpoint
// (this step's t0 is the previous step's t1, or - after a rejection - its t0), and
// _evaluate_leaf is a pure function of (key, r), so reusing the stored value is bit-identical
// to re-walking the tree.
template <class R, bool STLA>
struct BrownianTree {
  Key leaf;
  R T0, T1, sqrt_len;
  R memo_t[2];
  LevyVal<R> memo_v[2];
  VbtCache<R, STLA> cache;

  // the descent cache lives in dynamic shared memory: `smem` = the CTA's cache area, this thread uses column threadIdx.x
  __device__ __forceinline__ void attach_cache(R *smem, const VbtParams &vp) {
    cache.base = smem + threadIdx.x;
    cache.levels = vp.cache_levels;
    cache.stride = vp.cache_stride;
  }

  // leaf `w` of split_by_tree(key, shape) = jr.split(key, NUM)[w] (tree.py:301, _misc.py:128-133): NUM = 1 for shape=()
  template <int NUM>
  __device__ __forceinline__ void init_leaf(const uint32_t *user_key, int w, const VbtParams &vp) {
    init(user_key, vp);
    Key k{user_key[0], user_key[1]};
    leaf = split_child<NUM>(k, w, vp.partitionable != 0);
  }

  __device__ __forceinline__ void init(const uint32_t *user_key, const VbtParams &vp) {
    cache.n_entry = 0;
    cache.path = 0;
    Key k{user_key[0], user_key[1]};
    leaf = split_child<1>(k, 0, vp.partitionable != 0);  // split_by_tree(key, shape=()) == split(key, 1)[0]
    T0 = (R)vp.t0;
    T1 = (R)vp.t1;
    sqrt_len = r_sqrt(T1 - T0);
    memo_t[0] = memo_t[1] = Num<R>::nan();
  }

  __device__ __forceinline__ LevyVal<R> at(R t, const VbtParams &vp) {
    if (t == memo_t[1]) return memo_v[1];
    if (t == memo_t[0]) return memo_v[0];
    return vbt_evaluate_leaf<R, STLA>(leaf, linear_rescale(T0, t, T1), vp, &cache);
  }

  // evaluate(ta, tb, use_levy=True): tree.py:326-354 + _levy_diff + _denormalise_bm_inc
  __device__ __forceinline__ void increment(R ta, R tb, const VbtParams &vp, R &W, R &H) {
    const LevyVal<R> x0 = at(ta, vp);
    const LevyVal<R> x1 = at(tb, vp);
    memo_t[0] = ta; memo_v[0] = x0;
    memo_t[1] = tb; memo_v[1] = x1;
    const R su = x1.dt - x0.dt;
    const R w_su = x1.W - x0.W;
    W = sqrt_len * w_su;
    if constexpr (STLA) {
      const R inverse_su = R(1) / (r_abs(su) < Num<R>::eps() ? Num<R>::inf() : su);
      const R u_bb_s = x1.dt * x0.W - x0.dt * x1.W;
      const R bhh_su = x1.barH - x0.barH - R(0.5) * u_bb_s;
      H = sqrt_len * (inverse_su * bhh_su);
    } else {
      H = R(0);
    }
  }
};

}  // namespace dfx
