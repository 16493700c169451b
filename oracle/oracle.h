/* oracle.h - CPU restatement of Diffrax's ensemble hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may call into this library; the product (diffrax_b200/) never does.
 *
 * PARITY STATUS: "parity unpinned" at the north-star tolerances against a live
 * Diffrax: jax / equinox / optimistix are not importable in the authoring container
 *  and the reference ships no golden vectors for this path (§8c).
 * baseline/ holds the live-reference arm (probe + case builder + gen_golden.py): wherever jax + diffrax
 * are importable, tests/test_live_reference.py pins this oracle against Diffrax itself.
 * The oracle is pinned instead against every offline anchor the reference's tests
 * use: the Random123 threefry2x32 known-answer vectors, analytic solutions
 * (test_integrate.py:48-141, test_saveat_solution.py:29-195), scipy DOP853 on the
 * DETEST problems (test_detest.py:390-469), Butcher order conditions
 * (runge_kutta.py:126-130) and the statistical Brownian tests (test_brownian.py).
 *
 * Every function cites the reference file:line (under /root/reference/diffrax) it restates.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_DIM 128   /* (8 in the per-thread kernels; the warp-per-trajectory kernel of wide functors is checked up to 128) */
#define ORC_MAX_STAGES 14

/* dtype */
enum { ORC_F64 = 0, ORC_F32 = 1 };
/* solver ids (same numbering as tools/gen_tableaux.py) */
enum { ORC_TSIT5 = 0, ORC_DOPRI5 = 1, ORC_DOPRI8 = 2, ORC_HEUN = 3, ORC_BOSH3 = 4,
       ORC_MIDPOINT = 5, ORC_RALSTON = 6, ORC_EULER = 7, ORC_SHARK = 8,
       ORC_HALF = 0x100 /* flag: HalfSolver(inner), _solver/base.py:250-346 */ };
enum { ORC_EVENT_NONE = 0, ORC_EVENT_AFFINE = 1, ORC_EVENT_STEADY_STATE = 2 };
/* controller */
enum { ORC_CTRL_CONSTANT = 0, ORC_CTRL_PID = 1 };
/* vector fields */
enum { ORC_FIELD_DECAY = 0, ORC_FIELD_LOTKA_VOLTERRA = 1, ORC_FIELD_LORENZ = 2,
       ORC_FIELD_CR3BP = 3, ORC_FIELD_MLP = 4, ORC_FIELD_OU = 5, ORC_FIELD_FORCED_OSC = 6,
       ORC_FIELD_VDP = 7,
       ORC_FIELD_GBM = 8,        /* dy = mu y dt + sigma y dW: state-dependent diagonal diffusion, params [mu, sigma] */
       ORC_FIELD_OU_MATRIX = 16, /* + m: OU drift with a constant [d, m] diffusion matrix; params [theta, mu, G row-major] */
       ORC_FIELD_CALLBACK = 100 };
/* Brownian levy_area kind */
enum { ORC_LEVY_NONE = 0, ORC_LEVY_BROWNIAN_INCREMENT = 1, ORC_LEVY_SPACE_TIME = 2 };
/* RESULTS (_solution.py:13-31; only successful == 0 is pinned by the reference tests) */
enum { ORC_OK = 0, ORC_MAX_STEPS_REACHED = 1, ORC_DT_MIN_REACHED = 2, ORC_EVENT_OCCURRED = 3, ORC_EVENT_ROOT_FIND_FAILED = 4,
       ORC_MAX_STEPS_REJECTED = 5, ORC_INTERNAL_ERROR = 6 };

/* generic user vector field (ctypes callback): out[d] = f(t, y[d]) in double */
typedef void (*orc_callback_vf)(double t, const double *y, double *out, int dim);

typedef struct orc_desc {
  /* problem */
  int32_t field_id, dim, dtype, solver_id;
  const double *field_params; /* host doubles, meaning depends on field */
  int32_t n_field_params;
  orc_callback_vf callback;   /* ORC_FIELD_CALLBACK only (fp64 only) */
  /* batch + time */
  int64_t n_traj;
  const void *y0;            /* [N, d] REAL */
  double t0, t1;             /* used when the per-trajectory arrays are NULL */
  const void *t0_per_traj;   /* [N] REAL or NULL */
  const void *t1_per_traj;   /* [N] REAL or NULL */
  double dt0;                /* NaN => None */
  /* controller (pid.py:299-311) */
  int32_t controller;
  double rtol, atol, pcoeff, icoeff, dcoeff, safety, factormin, factormax;
  double dtmin, dtmax;       /* NaN => None */
  int32_t force_dtmin;
  double error_order;        /* NaN => solver.error_order(terms) */
  int32_t hairer_initial_step; /* 0 => dt0=None means 0.01; 1 => pid.py:51-81 */
  /* ClipStepSizeController(controller, step_ts, jump_ts) (clip.py:120-428); sorted ascending, user time */
  const void *step_ts; int32_t n_step_ts;
  const void *jump_ts; int32_t n_jump_ts;
  int32_t store_rejected_steps; /* 0 = None; else the length of the rejected-times stack (clip.py:292-299, 398-424) */
  /* SaveAt (_saveat.py:22-26,72-76) */
  int32_t save_t0, save_t1, save_steps, save_dense;
  const void *save_ts;       /* [T] REAL or NULL */
  int32_t n_save_ts;
  int32_t max_steps;
  /* outputs (caller-allocated, host) */
  int32_t out_size;          /* T_out, see orc_out_size */
  void *ts_out;              /* [N, T_out] */
  void *ys_out;              /* [N, T_out, d] */
  int32_t *stats;            /* [N, 3]: num_steps, accepted, rejected */
  int32_t *result;           /* [N] */
  void *dense_ts;            /* [N, max_steps+1] */
  void *dense_y0, *dense_y1; /* [N, max_steps, d] */
  void *dense_k;             /* [N, max_steps, s, d] (absent for 2-point interpolants) */
  int32_t *dense_count;      /* [N] number of accepted steps stored */
  void *y_final;             /* [N, d] optional */
  void *t_final;             /* [N] optional */
  /* Brownian motion (tree.py:245-301) */
  int32_t levy_area;
  const uint32_t *bm_keys;   /* [N, 2] user keys (before split_by_tree) */
  double bm_t0, bm_t1, bm_tol;
  int32_t threefry_partitionable;
  int32_t bm_dim;            /* 0: VirtualBrownianTree(shape=()); m > 0: shape=(m,) with m == dim - one independent tree per state
                              * component (leaf keys split_by_tree(key, (m,)) = split(key, m), tree.py:301, _misc.py:128-133) driving
                              * a diagonal diffusion */
  /* Event(cond_fn, root_finder, direction) (_event.py:13-118; _integrate.py:542-633, 691-821) with one registered condition:
   *   ORC_EVENT_AFFINE        c(t, y) = w . y + wt * t + b      event_params = [w[0..d), b, wt]   (real-valued: sign change)
   *   ORC_EVENT_STEADY_STATE  rms(f(t, y)) < atol + rtol * rms(y) event_params = [rtol, atol]      (boolean, _event.py:120-170)
   * t is the solver's (direction-normalised) time, as in the reference's call cond_fn(tprev, y, ...).
   * event_direction: 0 = None (any crossing), 1 = True (upcrossing), 2 = False (downcrossing).
   * event_root_find: 0 = root_finder None; 1 = Newton(event_rtol, event_atol) on the local interpolant, bracketed to the
   *   triggering step ([EXT] optimistix.Newton, restated from its published algorithm: clipped Newton steps from the step's
   *   end, Cauchy termination on both the iterate and the function value, at most 256 iterations). */
  int32_t n_events;                          /* 0 = no event; up to 4 conditions, the first that triggers wins (619-626) */
  int32_t event_kind[4], event_direction[4];
  int32_t event_root_find;
  const double *event_params; int32_t n_event_params; /* the conditions' parameters back to back */
  double event_rtol, event_atol;
  /* Resuming: diffeqsolve(..., solver_state=, controller_state=, made_jump=) and SaveAt(solver_state=True, ...)
   * (_integrate.py:1250-1271, 1489-1500).  One record per trajectory, [N, 5 + d] REAL:
   *   [0] prev_inv_scaled_error  [1] prev_prev_inv_scaled_error  [2] at_dtmin   (PIDController state, pid.py:388-392)
   *   [3] made_jump   [4] first_step  [5 ..] the carried FSAL derivative        (solver state, runge_kutta.py:415-444)
   * state_in_flags: bit 0 controller_state passed, bit 1 solver_state passed, bit 2 made_jump passed. */
  const void *state_in; int32_t state_in_flags;
  void *state_out;
  /* test hook: per-step trace of (tprev, tnext, keep) for trajectory `trace_traj` */
  int64_t trace_traj;
  double *trace;             /* [max_steps, 3] or NULL */
  int32_t num_threads;       /* worker threads, 0 => all online cores */
} orc_desc;

int orc_out_size(const orc_desc *d);           /* _integrate.py:1273-1293 */
int orc_solve(const orc_desc *d);              /* 0 ok; <0 argument error */
const char *orc_last_error(void);
int orc_num_stages(int solver_id);
int orc_hw_threads(void);

/* PRNG (jax/_src/prng.py) */
void orc_threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t *o0, uint32_t *o1);
void orc_split(uint32_t k0, uint32_t k1, int num, int partitionable, uint32_t *out /* [num,2] */);
double orc_normal_f64(uint32_t k0, uint32_t k1, int partitionable);
float orc_normal_f32(uint32_t k0, uint32_t k1, int partitionable);
double orc_erfinv_f64(double x);
float orc_erfinv_f32(float x);
/* element w of jr.normal(key, (m,)); m == 1, w == 0 is the scalar draw */
double orc_normal_vec_f64(uint32_t k0, uint32_t k1, int w, int m, int partitionable);
float orc_normal_vec_f32(uint32_t k0, uint32_t k1, int w, int m, int partitionable);
/* the explicitly sequenced log1p that erf_inv is built on (see oracle.c) */
double orc_log1p_f64(double x);
float orc_log1p_f32(float x);

/* VirtualBrownianTree.evaluate(t0, t1, use_levy=True) for n independent trees of shape () (bm_dim == 0)
 * or (bm_dim,) (tree.py:326-354).  keys: [n,2] user keys; outputs W[n(, m)], H[n(, m)] (H may be NULL). */
int orc_vbt_evaluate(int dtype, int levy_area, int partitionable, int64_t n, const uint32_t *keys,
                     double bm_t0, double bm_t1, double bm_tol, const void *ta, const void *tb,
                     int per_traj_times, void *W, void *H, int bm_dim);

/* DenseInterpolation.evaluate (_global_interpolation.py:335-355) for one batch of queries:
 * each trajectory i evaluates at tq[i*nq + q]. out: [N, nq, d] */
int orc_dense_evaluate(int dtype, int solver_id, int64_t n_traj, int dim, int max_steps,
                       const void *dense_ts, const void *dense_y0, const void *dense_y1,
                       const void *dense_k, const int32_t *dense_count, double direction,
                       const void *tq, int nq, void *out);
/* DenseInterpolation.derivative, _global_interpolation.py:357-368 */
int orc_dense_derivative(int dtype, int solver_id, int64_t n_traj, int dim, int max_steps,
                       const void *dense_ts, const void *dense_y0, const void *dense_y1,
                       const void *dense_k, const int32_t *dense_count, double direction,
                       const void *tq, int nq, void *out);

#ifdef __cplusplus
}
#endif
#endif
