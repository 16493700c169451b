/* oracle.c - CPU restatement of Diffrax's vmapped diffeqsolve hot path.
 * TEST INFRASTRUCTURE ONLY - see oracle.h for who may call this and for the parity status
 * ("parity unpinned" against a live Diffrax; pinned to the reference's offline anchors).
 *
 * Build: see oracle/Makefile.  Compiled with -ffp-contract=off so that every a*b+c is the
 * two-rounding IEEE result the expression order in the reference denotes; the product's
 * CUDA kernels are free to contract to FMA (as XLA's LLVM backends also do), which is why
 * the parity tests compare at the north-star tolerances rather than bit-for-bit.
 */
#include "oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#include "oracle_tableaux.h"

static char g_err[256] = "";
const char *orc_last_error(void) { return g_err; }

/* ---------------------------------------------------------------------------------------
 * threefry2x32, 20 rounds (Salmon et al. 2011; jax/_src/prng.py `threefry2x32` -
 * [EXT: not under /root/reference; pinned by the Random123 known-answer vectors, SURVEY §8c])
 * ------------------------------------------------------------------------------------- */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

void orc_threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t *o0, uint32_t *o1) {
  static const int R0[4] = {13, 15, 26, 6}, R1[4] = {17, 29, 16, 24};
  uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  x0 += ks[0];
  x1 += ks[1];
  for (int g = 1; g <= 5; ++g) {
    const int *rot = (g & 1) ? R0 : R1;
    for (int i = 0; i < 4; ++i) {
      x0 += x1;
      x1 = rotl32(x1, rot[i]);
      x1 ^= x0;
    }
    x0 += ks[g % 3];
    x1 += ks[(g + 1) % 3] + (uint32_t)g;
  }
  *o0 = x0;
  *o1 = x1;
}

/* jax.random.split(key, num) (prng.py `_threefry_split`).
 *  partitionable ("foldlike"): out[i] = threefry(key, (0, i))
 *  original: counts = iota(2*num); (x0s, x1s) = halves; out flat = concat(y0s, y1s), reshaped (num, 2) */
void orc_split(uint32_t k0, uint32_t k1, int num, int partitionable, uint32_t *out) {
  if (partitionable) {
    for (int i = 0; i < num; ++i) orc_threefry2x32(k0, k1, 0u, (uint32_t)i, &out[2 * i], &out[2 * i + 1]);
  } else {
    uint32_t *flat = out; /* flat[0..num) = y0s, flat[num..2num) = y1s */
    for (int i = 0; i < num; ++i) {
      uint32_t a, b;
      orc_threefry2x32(k0, k1, (uint32_t)i, (uint32_t)(num + i), &a, &b);
      flat[i] = a;
      flat[num + i] = b;
    }
  }
}

/* ---------------------------------------------------------------------------------------
 * The float side of jax.random.normal.  [EXT - jax / XLA, not under /root/reference.]
 *
 * XLA evaluates normal = sqrt(2) * erf_inv(u) with erf_inv = Giles' (2010) polynomials in
 * w = -log1p(-u*u) (xla/client/lib/math.cc ErfInv32 / ErfInv64).  The last ulp of that chain
 * is backend-specific in XLA itself (log1p comes from the backend's libm / libdevice and the
 * LLVM back ends may contract a*b+c), so there is no single "JAX bit pattern" to restate.
 * What this project pins instead is ONE explicitly sequenced evaluation, stated operation by
 * operation below, that the CUDA side (csrc/prng.cuh) performs with the identical sequence of
 * IEEE-754 correctly rounded operations (+, -, *, /, sqrt, fma): oracle and kernel agree bit
 * for bit, and the sequence is within 1 ulp of the exact log1p (tests/test_oracle_prng.py).
 *   - every "fma(a, b, c)" below is ONE rounding; every other operator rounds on its own
 *     (this file is compiled with -ffp-contract=off);
 *   - log1p (needed on [-1, 0] only) follows the classical argument reduction 1+x = 2^k (1+f),
 *     f in [sqrt(1/2)-1, sqrt(2)-1), log(1+f) = f - f^2/2 + s (f^2/2 + R(s^2)), s = f / (2 + f) (the published
 *     fdlibm scheme and coefficients), branch-free: f comes from the exact x by one fma, so 1+x's rounding never enters.
 * ------------------------------------------------------------------------------------- */
double orc_log1p_f64(double x) { /* domain: -1 <= x <= 0 (the only one erf_inv needs: x = -u*u) */
  static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                      Lp[7] = {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
                               2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
                               1.479819860511658591e-01};
  if (!(x > -1.0 && x <= 0.0)) return x == -1.0 ? -INFINITY : NAN;
  /* k from the (rounded) 1 + x; for x > sqrt(1/2) - 1 no reduction is needed */
  const double u = 1.0 + x;
  uint64_t b;
  memcpy(&b, &u, 8);
  int k = (int)((b >> 52) & 0x7ff) - 1023 + ((b & 0x000fffffffffffffull) >= 0x6a09e667f3bcdull ? 1 : 0);
  if (x > -0.2928932188134524) k = 0;
  /* 1 + f = 2^-k (1 + x)  =>  f = 2^-k x + (2^-k - 1): ONE fma from the exact x, so the rounding of 1 + x never enters */
  const uint64_t sb = (uint64_t)(1023 - k) << 52;
  double scale;
  memcpy(&scale, &sb, 8);
  const double f = fma(scale, x, scale - 1.0);
  const double hfsq = (0.5 * f) * f;
  const double s = f / (2.0 + f);
  const double z = s * s;
  double r = Lp[6];
  for (int i = 5; i >= 0; --i) r = fma(r, z, Lp[i]);
  r = r * z;
  const double dk = (double)k;
  return dk * ln2_hi - ((hfsq - (s * (hfsq + r) + dk * ln2_lo)) - f);
}
float orc_log1p_f32(float x) {
  static const float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f,
                     Lp[7] = {6.6666668653e-01f, 4.0000000596e-01f, 2.8571429849e-01f, 2.2222198546e-01f,
                              1.8183572590e-01f, 1.5313838422e-01f, 1.4798198640e-01f};
  if (!(x > -1.0f && x <= 0.0f)) return x == -1.0f ? -INFINITY : NAN;
  const float u = 1.0f + x;
  uint32_t b;
  memcpy(&b, &u, 4);
  int k = (int)((b >> 23) & 0xff) - 127 + ((b & 0x007fffffu) >= 0x3504f3u ? 1 : 0);
  if (x > -0.29289323f) k = 0;
  const uint32_t sb = (uint32_t)(127 - k) << 23;
  float scale;
  memcpy(&scale, &sb, 4);
  const float f = fmaf(scale, x, scale - 1.0f);
  const float hfsq = (0.5f * f) * f;
  const float s = f / (2.0f + f);
  const float z = s * s;
  float r = Lp[6];
  for (int i = 5; i >= 0; --i) r = fmaf(r, z, Lp[i]);
  r = r * z;
  const float dk = (float)k;
  return dk * ln2_hi - ((hfsq - (s * (hfsq + r) + dk * ln2_lo)) - f);
}

/* lax.erf_inv, f32: Giles (2010) single-precision polynomial, coefficients and branch as XLA's ErfInv32
 * [EXT].  Sequence: w = -log1p(-(x*x)); Horner with one fma per coefficient; result p * x. */
float orc_erfinv_f32(float x) {
  static const float lt5[9] = {2.81022636e-08f, 3.43273939e-07f, -3.5233877e-06f, -4.39150654e-06f, 0.00021858087f,
                               -0.00125372503f, -0.00417768164f, 0.246640727f, 1.50140941f};
  static const float ge5[9] = {-0.000200214257f, 0.000100950558f, 0.00134934322f, -0.00367342844f, 0.00573950773f,
                               -0.0076224613f, 0.00943887047f, 1.00167406f, 2.83297682f};
  if (fabsf(x) == 1.0f) return x * INFINITY;
  float w = -orc_log1p_f32(-(x * x));
  const float *c;
  if (w < 5.0f) { w = w - 2.5f; c = lt5; } else { w = sqrtf(w) - 3.0f; c = ge5; }
  float p = c[0];
  for (int i = 1; i < 9; ++i) p = fmaf(p, w, c[i]);
  return p * x;
}

/* lax.erf_inv, f64: Giles' double-precision three-branch polynomial, coefficients and branches as XLA's
 * ErfInv64 [EXT]; validated against scipy.special.erfinv in tests/test_oracle_prng.py.  Same sequencing. */
double orc_erfinv_f64(double x) {
  static const double a[23] = {-3.6444120640178196996e-21, -1.685059138182016589e-19, 1.2858480715256400167e-18,
                               1.115787767802518096e-17, -1.333171662854620906e-16, 2.0972767875968561637e-17,
                               6.6376381343583238325e-15, -4.0545662729752068639e-14, -8.1519341976054721522e-14,
                               2.6335093153082322977e-12, -1.2975133253453532498e-11, -5.4154120542946279317e-11,
                               1.051212273321532285e-09, -4.1126339803469836976e-09, -2.9070369957882005086e-08,
                               4.2347877827932403518e-07, -1.3654692000834678645e-06, -1.3882523362786468719e-05,
                               0.0001867342080340571352, -0.00074070253416626697512, -0.0060336708714301490533,
                               0.24015818242558961693, 1.6536545626831027356};
  static const double b[19] = {2.2137376921775787049e-09, 9.0756561938885390979e-08, -2.7517406297064545428e-07,
                               1.8239629214389227755e-08, 1.5027403968909827627e-06, -4.013867526981545969e-06,
                               2.9234449089955446044e-06, 1.2475304481671778723e-05, -4.7318229009055733981e-05,
                               6.8284851459573175448e-05, 2.4031110387097893999e-05, -0.0003550375203628474796,
                               0.00095328937973738049703, -0.0016882755560235047313, 0.0024914420961078508066,
                               -0.0037512085075692412107, 0.005370914553590063617, 1.0052589676941592334,
                               3.0838856104922207635};
  static const double c[17] = {-2.7109920616438573243e-11, -2.5556418169965252055e-10, 1.5076572693500548083e-09,
                               -3.7894654401267369937e-09, 7.6157012080783393804e-09, -1.4960026627149240478e-08,
                               2.9147953450901080826e-08, -6.7711997758452339498e-08, 2.2900482228026654717e-07,
                               -9.9298272942317002539e-07, 4.5260625972231537039e-06, -1.9681778105531670567e-05,
                               7.5995277030017761139e-05, -0.00021503011930044477347, -0.00013871931833623122026,
                               1.0103004648645343977, 4.8499064014085844221};
  if (fabs(x) == 1.0) return x * INFINITY;
  double w = -orc_log1p_f64(-(x * x));
  const double *co;
  int n;
  if (w < 6.25) { w = w - 3.125; co = a; n = 23; }
  else if (w < 16.0) { w = sqrt(w) - 3.25; co = b; n = 19; }
  else { w = sqrt(w) - 5.0; co = c; n = 17; }
  double p = co[0];
  for (int i = 1; i < n; ++i) p = fma(p, w, co[i]);
  return p * x;
}

/* random bits of element `w` of a draw of shape (m,) (prng.py threefry_random_bits; m = 1 is shape ()):
 *  partitionable: counter (0, w): 32-bit -> x0 ^ x1 ; 64-bit -> (x0 << 32) | x1
 *  original 32-bit: counters iota(m) padded with one 0 to even length 2h, halves (c[:h], c[h:]); the words are
 *                   concat(x0s, x1s)[:m], so element w < h is x0 of block (w, h+w or 0 for the pad) and element
 *                   w >= h is x1 of block (w-h, w)
 *  original 64-bit: 2m 32-bit words from counters iota(2m), halves (c[:m], c[m:]); high words = x0s, low = x1s,
 *                   so element w is (x0 << 32) | x1 of block (w, m+w) */
static uint32_t bits32_vec(uint32_t k0, uint32_t k1, int w, int m, int partitionable) {
  uint32_t a, b;
  if (partitionable) { orc_threefry2x32(k0, k1, 0u, (uint32_t)w, &a, &b); return a ^ b; }
  const int h = (m + 1) / 2;
  if (w < h) { orc_threefry2x32(k0, k1, (uint32_t)w, (h + w < m) ? (uint32_t)(h + w) : 0u, &a, &b); return a; }
  orc_threefry2x32(k0, k1, (uint32_t)(w - h), (uint32_t)w, &a, &b);
  return b;
}
static uint64_t bits64_vec(uint32_t k0, uint32_t k1, int w, int m, int partitionable) {
  uint32_t a, b;
  if (partitionable) orc_threefry2x32(k0, k1, 0u, (uint32_t)w, &a, &b);
  else orc_threefry2x32(k0, k1, (uint32_t)w, (uint32_t)(m + w), &a, &b);
  return ((uint64_t)a << 32) | (uint64_t)b;
}

/* jax.random.normal(key, shape, dtype)[w] (random.py `_normal_real` / `_uniform`):
 * mantissa fill -> [1,2) -> -1 -> u = max(lo, fma(f, hi - lo, lo)), lo = nextafter(-1, 0), hi = 1;
 * normal = sqrt(2) * erf_inv(u). */
float orc_normal_vec_f32(uint32_t k0, uint32_t k1, int w, int m, int partitionable) {
  uint32_t bits = bits32_vec(k0, k1, w, m, partitionable);
  uint32_t fb = (bits >> 9) | 0x3F800000u;
  float f;
  memcpy(&f, &fb, 4);
  f = f - 1.0f;
  const float lo = nextafterf(-1.0f, 0.0f), hi = 1.0f;
  float u = fmaf(f, hi - lo, lo);
  if (!(u > lo)) u = lo; /* lax.max(lo, u) */
  return (float)sqrt(2.0) * orc_erfinv_f32(u);
}
double orc_normal_vec_f64(uint32_t k0, uint32_t k1, int w, int m, int partitionable) {
  uint64_t bits = bits64_vec(k0, k1, w, m, partitionable);
  uint64_t fb = (bits >> 12) | 0x3FF0000000000000ull;
  double f;
  memcpy(&f, &fb, 8);
  f = f - 1.0;
  const double lo = nextafter(-1.0, 0.0), hi = 1.0;
  double u = fma(f, hi - lo, lo);
  if (!(u > lo)) u = lo;
  return sqrt(2.0) * orc_erfinv_f64(u);
}
float orc_normal_f32(uint32_t k0, uint32_t k1, int partitionable) { return orc_normal_vec_f32(k0, k1, 0, 1, partitionable); }
double orc_normal_f64(uint32_t k0, uint32_t k1, int partitionable) { return orc_normal_vec_f64(k0, k1, 0, 1, partitionable); }

/* ---------------------------------------------------------------------------------------
 * type-generic core, instantiated for f64 and f32
 * ------------------------------------------------------------------------------------- */
#define REAL double
#define FN(x) x##_f64
#define R(x) ((double)(x))
#include "oracle_core.inc"
#undef REAL
#undef FN
#undef R

#define REAL float
#define FN(x) x##_f32
#define R(x) ((float)(x))
#include "oracle_core.inc"
#undef REAL
#undef FN
#undef R


/* ---------------------------------------------------------------------------------------
 * tiny pthread work-sharing loop (dynamic chunks of 64 trajectories).  OpenMP is not used:
 * this image's gcc driver cannot find libgomp.spec.
 * ------------------------------------------------------------------------------------- */
typedef void (*range_fn)(void *ctx, int64_t lo, int64_t hi);
typedef struct { atomic_llong next; int64_t n; range_fn fn; void *ctx; } pf_shared;
static void *pf_worker(void *arg) {
  pf_shared *sh = (pf_shared *)arg;
  for (;;) {
    long long lo = atomic_fetch_add(&sh->next, 64);
    if (lo >= sh->n) break;
    long long hi = lo + 64 < sh->n ? lo + 64 : sh->n;
    sh->fn(sh->ctx, lo, hi);
  }
  return NULL;
}
int orc_hw_threads(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }
static void parallel_for(int64_t n, int num_threads, range_fn fn, void *ctx) {
  int nt = num_threads > 0 ? num_threads : orc_hw_threads();
  if (nt > 256) nt = 256;
  if ((int64_t)nt * 64 > n) nt = (int)((n + 63) / 64);
  if (nt <= 1) { fn(ctx, 0, n); return; }
  pf_shared sh; atomic_init(&sh.next, 0); sh.n = n; sh.fn = fn; sh.ctx = ctx;
  pthread_t th[256];
  for (int i = 1; i < nt; ++i) pthread_create(&th[i], NULL, pf_worker, &sh);
  pf_worker(&sh);
  for (int i = 1; i < nt; ++i) pthread_join(th[i], NULL);
}
static void solve_range(void *ctx, int64_t lo, int64_t hi) {
  const orc_desc *d = (const orc_desc *)ctx;
  for (int64_t i = lo; i < hi; ++i) {
    if (d->dtype == ORC_F64) solve_one_f64(d, i);
    else solve_one_f32(d, i);
  }
}

int orc_num_stages(int solver_id) {
  solver_id &= ~ORC_HALF; /* HalfSolver(inner): the inner solver's stages / interpolant */
  for (int i = 0; i < ORC_NUM_TABLEAUX; ++i)
    if (orc_tableaux[i].id == solver_id) return orc_tableaux[i].stages;
  if (solver_id == ORC_SHARK) return 2;
  if (solver_id == ORC_EULER) return 1;
  return -1;
}

/* _integrate.py:1273-1293 _allocate_output */
int orc_out_size(const orc_desc *d) {
  int out = 0;
  if (d->save_t0) out += 1;
  if (d->save_ts) out += d->n_save_ts;
  if (d->save_steps != 0) out += d->max_steps / d->save_steps;
  if (d->save_t1 && (d->save_steps == 0 || (d->max_steps % d->save_steps) != 0)) out += 1;
  return out;
}

int orc_solve(const orc_desc *d) {
  if (d->dim < 1 || d->dim > ORC_MAX_DIM) { snprintf(g_err, sizeof g_err, "dim %d out of range", d->dim); return -1; }
  if (orc_num_stages(d->solver_id) < 0) { snprintf(g_err, sizeof g_err, "unknown solver %d", d->solver_id); return -1; }
  if (d->out_size != orc_out_size(d)) { snprintf(g_err, sizeof g_err, "out_size mismatch"); return -1; }
  if (d->controller == ORC_CTRL_CONSTANT && d->dt0 != d->dt0) { snprintf(g_err, sizeof g_err, "constant steps need dt0"); return -1; }
  if (d->levy_area != ORC_LEVY_NONE && !d->bm_keys) { snprintf(g_err, sizeof g_err, "SDE needs keys"); return -1; }
  parallel_for(d->n_traj, d->num_threads, solve_range, (void *)d);
  return 0;
}

int orc_vbt_evaluate(int dtype, int levy_area, int partitionable, int64_t n, const uint32_t *keys, double bm_t0,
                     double bm_t1, double bm_tol, const void *ta, const void *tb, int per_traj_times, void *W, void *H,
                     int bm_dim) {
  const int m = bm_dim > 0 ? bm_dim : 1;
  if (m > ORC_MAX_DIM) { snprintf(g_err, sizeof g_err, "bm_dim %d out of range", bm_dim); return -1; }
  for (int64_t i = 0; i < n; ++i) {
    if (dtype == ORC_F64) {
      double a = ((const double *)ta)[per_traj_times ? i : 0], b = ((const double *)tb)[per_traj_times ? i : 0];
      double w[ORC_MAX_DIM], h[ORC_MAX_DIM];
      vbt_increment_vec_f64(keys + 2 * i, m, bm_t0, bm_t1, bm_tol, levy_area, partitionable, a, b, w, h);
      for (int c = 0; c < m; ++c) { ((double *)W)[i * m + c] = w[c]; if (H) ((double *)H)[i * m + c] = h[c]; }
    } else {
      float a = ((const float *)ta)[per_traj_times ? i : 0], b = ((const float *)tb)[per_traj_times ? i : 0];
      float w[ORC_MAX_DIM], h[ORC_MAX_DIM];
      vbt_increment_vec_f32(keys + 2 * i, m, bm_t0, bm_t1, bm_tol, levy_area, partitionable, a, b, w, h);
      for (int c = 0; c < m; ++c) { ((float *)W)[i * m + c] = w[c]; if (H) ((float *)H)[i * m + c] = h[c]; }
    }
  }
  return 0;
}

static int dense_eval_entry(int deriv, int dtype, int solver_id, int64_t n_traj, int dim, int max_steps, const void *dense_ts,
                       const void *dense_y0, const void *dense_y1, const void *dense_k, const int32_t *dense_count,
                       double direction, const void *tq, int nq, void *out) {
  solver_id &= ~ORC_HALF;
  const int s = orc_num_stages(solver_id);
  for (int64_t i = 0; i < n_traj; ++i) {
    for (int q = 0; q < nq; ++q) {
      if (dtype == ORC_F64) {
        dense_eval_one_f64(solver_id, dim, max_steps, (const double *)dense_ts + (size_t)i * (max_steps + 1),
                           (const double *)dense_y0 + (size_t)i * max_steps * dim,
                           (const double *)dense_y1 + (size_t)i * max_steps * dim,
                           dense_k ? (const double *)dense_k + (size_t)i * max_steps * s * dim : NULL, dense_count[i],
                           direction, ((const double *)tq)[(size_t)i * nq + q], deriv, (double *)out + ((size_t)i * nq + q) * dim);
      } else {
        dense_eval_one_f32(solver_id, dim, max_steps, (const float *)dense_ts + (size_t)i * (max_steps + 1),
                           (const float *)dense_y0 + (size_t)i * max_steps * dim,
                           (const float *)dense_y1 + (size_t)i * max_steps * dim,
                           dense_k ? (const float *)dense_k + (size_t)i * max_steps * s * dim : NULL, dense_count[i],
                           (float)direction, ((const float *)tq)[(size_t)i * nq + q], deriv, (float *)out + ((size_t)i * nq + q) * dim);
      }
    }
  }
  return 0;
}
int orc_dense_evaluate(int dtype, int solver_id, int64_t n_traj, int dim, int max_steps, const void *dense_ts,
                       const void *dense_y0, const void *dense_y1, const void *dense_k, const int32_t *dense_count,
                       double direction, const void *tq, int nq, void *out) {
  return dense_eval_entry(0, dtype, solver_id, n_traj, dim, max_steps, dense_ts, dense_y0, dense_y1, dense_k, dense_count, direction, tq, nq, out);
}
int orc_dense_derivative(int dtype, int solver_id, int64_t n_traj, int dim, int max_steps, const void *dense_ts,
                         const void *dense_y0, const void *dense_y1, const void *dense_k, const int32_t *dense_count,
                         double direction, const void *tq, int nq, void *out) {
  return dense_eval_entry(1, dtype, solver_id, n_traj, dim, max_steps, dense_ts, dense_y0, dense_y1, dense_k, dense_count, direction, tq, nq, out);
}
