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

/* random_bits for a scalar (shape == ()) draw (prng.py `threefry_random_bits`):
 *  partitionable: block (0,0): 32-bit -> x0 ^ x1 ; 64-bit -> (x0 << 32) | x1
 *  original:      32-bit -> first word of block (0,0) ; 64-bit -> block (0,1): (x0 << 32) | x1 */
static uint32_t bits32(uint32_t k0, uint32_t k1, int partitionable) {
  uint32_t a, b;
  orc_threefry2x32(k0, k1, 0u, 0u, &a, &b);
  return partitionable ? (a ^ b) : a;
}
static uint64_t bits64(uint32_t k0, uint32_t k1, int partitionable) {
  uint32_t a, b;
  orc_threefry2x32(k0, k1, 0u, partitionable ? 0u : 1u, &a, &b);
  return ((uint64_t)a << 32) | (uint64_t)b;
}

/* lax.erf_inv, f32: Giles (2010) single-precision polynomial as used by XLA (ErfInv32)
 * [EXT - restated from the published algorithm; not under /root/reference]. */
float orc_erfinv_f32(float x) {
  static const float lt5[9] = {2.81022636e-08f, 3.43273939e-07f, -3.5233877e-06f, -4.39150654e-06f, 0.00021858087f,
                               -0.00125372503f, -0.00417768164f, 0.246640727f, 1.50140941f};
  static const float ge5[9] = {-0.000200214257f, 0.000100950558f, 0.00134934322f, -0.00367342844f, 0.00573950773f,
                               -0.0076224613f, 0.00943887047f, 1.00167406f, 2.83297682f};
  if (fabsf(x) == 1.0f) return x * INFINITY;
  float w = -log1pf(-x * x);
  const float *c;
  if (w < 5.0f) { w = w - 2.5f; c = lt5; } else { w = sqrtf(w) - 3.0f; c = ge5; }
  float p = c[0];
  for (int i = 1; i < 9; ++i) p = c[i] + p * w;
  return p * x;
}

/* lax.erf_inv, f64: Giles' double-precision three-branch polynomial as used by XLA (ErfInv64)
 * [EXT - restated from the published algorithm; validated against scipy.special.erfinv in
 * tests/test_oracle_prng.py]. */
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
  double w = -log1p(-x * x);
  const double *co;
  int n;
  if (w < 6.25) { w = w - 3.125; co = a; n = 23; }
  else if (w < 16.0) { w = sqrt(w) - 3.25; co = b; n = 19; }
  else { w = sqrt(w) - 5.0; co = c; n = 17; }
  double p = co[0];
  for (int i = 1; i < n; ++i) p = co[i] + p * w;
  return p * x;
}

/* jax.random.normal(key, (), dtype) (random.py `_normal_real` / `_uniform`):
 * mantissa fill -> [1,2) -> -1 -> u = max(lo, f*(hi-lo)+lo), lo = nextafter(-1, 0), hi = 1;
 * normal = sqrt(2) * erf_inv(u). */
float orc_normal_f32(uint32_t k0, uint32_t k1, int partitionable) {
  uint32_t bits = bits32(k0, k1, partitionable);
  uint32_t fb = (bits >> 9) | 0x3F800000u;
  float f;
  memcpy(&f, &fb, 4);
  f = f - 1.0f;
  const float lo = nextafterf(-1.0f, 0.0f), hi = 1.0f;
  float u = f * (hi - lo) + lo;
  if (!(u > lo)) u = lo; /* lax.max(lo, u) */
  return (float)sqrt(2.0) * orc_erfinv_f32(u);
}
double orc_normal_f64(uint32_t k0, uint32_t k1, int partitionable) {
  uint64_t bits = bits64(k0, k1, partitionable);
  uint64_t fb = (bits >> 12) | 0x3FF0000000000000ull;
  double f;
  memcpy(&f, &fb, 8);
  f = f - 1.0;
  const double lo = nextafter(-1.0, 0.0), hi = 1.0;
  double u = f * (hi - lo) + lo;
  if (!(u > lo)) u = lo;
  return sqrt(2.0) * orc_erfinv_f64(u);
}

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
                     double bm_t1, double bm_tol, const void *ta, const void *tb, int per_traj_times, void *W, void *H) {
  for (int64_t i = 0; i < n; ++i) {
    if (dtype == ORC_F64) {
      double a = ((const double *)ta)[per_traj_times ? i : 0], b = ((const double *)tb)[per_traj_times ? i : 0];
      double w, h;
      vbt_increment_f64(keys + 2 * i, bm_t0, bm_t1, bm_tol, levy_area, partitionable, a, b, &w, &h);
      ((double *)W)[i] = w;
      if (H) ((double *)H)[i] = h;
    } else {
      float a = ((const float *)ta)[per_traj_times ? i : 0], b = ((const float *)tb)[per_traj_times ? i : 0];
      float w, h;
      vbt_increment_f32(keys + 2 * i, bm_t0, bm_t1, bm_tol, levy_area, partitionable, a, b, &w, &h);
      ((float *)W)[i] = w;
      if (H) ((float *)H)[i] = h;
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
