/* diffrax_b200.h - C ABI of the B200-native ensemble integrator.
 *
 * The reference (patrick-kidger/diffrax, pure Python on JAX) has NO native boundary for
 * this path: the only entry is the Python call
 *     diffeqsolve(terms, solver, t0, t1, dt0, y0, args, *, saveat, stepsize_controller,
 *                 max_steps, throw, ...) -> Solution            (diffrax/_integrate.py:890-911)
 * invoked under jax.vmap (test/test_vmap.py:28-39).  BASELINE.json's north star prescribes
 * the boundary: a jax.ffi custom call behind that same Python signature.  Each entry point
 * below is what such an FFI handler binds; the file:line beside it is the reference code
 * whose work it replaces.  INTEGRATION.md shows the jax.ffi / ctypes stubs.
 *
 * Conventions
 *  - plain C types only; every buffer is caller-owned; device entry points take device
 *    pointers and a cudaStream_t (as void*), enqueue on that stream and never synchronise
 *    the device; `_host` entry points take host pointers and do the H2D / D2H themselves.
 *  - reentrant: no mutable globals except the launcher registry (guarded) and the
 *    thread-local error string.
 *  - return 0 on success, a negative DFX_ERR_* on argument / launch errors.  Numerical
 *    failures are per-trajectory `result` codes (diffrax/_solution.py:13-31).
 *  - layouts are C-contiguous with the trajectory index first, i.e. exactly what
 *    jax.vmap(diffeqsolve) returns: ts[N,T], ys[N,T,d], stats[N,3].
 */
#ifndef DIFFRAX_B200_H
#define DIFFRAX_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFX_ABI_VERSION 2

enum dfx_dtype { DFX_F64 = 0, DFX_F32 = 1 };

/* Solvers: diffrax/_solver/{tsit5,dopri5,dopri8,heun,bosh3,midpoint,ralston,euler,shark}.py */
enum dfx_solver { DFX_TSIT5 = 0, DFX_DOPRI5 = 1, DFX_DOPRI8 = 2, DFX_HEUN = 3, DFX_BOSH3 = 4,
                  DFX_MIDPOINT = 5, DFX_RALSTON = 6, DFX_EULER = 7, DFX_SHARK = 8,
                  /* flag: HalfSolver(inner) is DFX_HALF_SOLVER | inner (diffrax/_solver/base.py:250-346) */
                  DFX_HALF_SOLVER = 0x100 };

/* Step size controllers: _step_size_controller/constant.py:20-104, pid.py:299-567 */
enum dfx_controller { DFX_CTRL_CONSTANT = 0, DFX_CTRL_PID = 1 };

/* Built-in registered vector-field functors (diffrax_b200/csrc/fields.cuh).
 * User functors register further ids >= DFX_FIELD_USER with dfx_register_launcher. */
enum dfx_field { DFX_FIELD_DECAY = 0, DFX_FIELD_LOTKA_VOLTERRA = 1, DFX_FIELD_LORENZ = 2,
                 DFX_FIELD_CR3BP = 3, DFX_FIELD_MLP = 4, DFX_FIELD_OU = 5,
                 DFX_FIELD_FORCED_OSC = 6, DFX_FIELD_VDP = 7,
                 DFX_FIELD_GBM = 8,        /* geometric Brownian motion: dy = mu y dt + sigma y dW (state-dependent diffusion) */
                 DFX_FIELD_OU_MATRIX = 16, /* + m (1..4): OU drift with a constant [d, m] diffusion matrix, params [theta, mu, G] */
                 DFX_FIELD_USER = 1000 };

/* VirtualBrownianTree(levy_area=...) (_brownian/tree.py:238-254) */
enum dfx_levy { DFX_LEVY_NONE = 0, DFX_LEVY_BROWNIAN_INCREMENT = 1, DFX_LEVY_SPACE_TIME = 2 };

/* RESULTS (_solution.py:13-31).  successful == 0 is the only value the reference pins
 * (test/test_saveat_solution.py:21). */
enum dfx_result { DFX_RESULT_SUCCESSFUL = 0, DFX_RESULT_MAX_STEPS_REACHED = 1,
                  DFX_RESULT_DT_MIN_REACHED = 2, DFX_RESULT_EVENT_OCCURRED = 3 /* not a failure: _solution.py:52-62 is_okay */,
                  DFX_RESULT_EVENT_ROOT_FIND_FAILED = 4, DFX_RESULT_MAX_STEPS_REJECTED = 5 /* _solution.py:24-27 */,
                  DFX_RESULT_INTERNAL_ERROR = 6 };
enum dfx_event { DFX_EVENT_NONE = 0, DFX_EVENT_AFFINE = 1, DFX_EVENT_STEADY_STATE = 2, DFX_EVENT_USER = 3 };
#define DFX_MAX_EVENTS 4
#define DFX_MAX_PEERS 8

enum dfx_error { DFX_OK = 0, DFX_ERR_BAD_ARGUMENT = -1, DFX_ERR_UNSUPPORTED = -2,
                 DFX_ERR_CUDA = -3, DFX_ERR_NO_DEVICE = -4 };

/* One vmapped diffeqsolve call.  Field-by-field mirror of the reference arguments. */
typedef struct dfx_solve_desc {
  uint32_t struct_size;       /* sizeof(dfx_solve_desc), for ABI evolution */
  uint32_t abi_version;       /* DFX_ABI_VERSION */

  /* terms: ODETerm(field) or MultiTerm(ODETerm(drift), ControlTerm(diffusion, VBT))
   * (_term.py:174-226, 271-555, 666-731) */
  int32_t field_id;           /* registered functor */
  int32_t dim;                /* state dimension d */
  int32_t dtype;              /* dfx_dtype: state == time dtype (_integrate.py:1060-1099) */
  int32_t solver_id;          /* dfx_solver */
  const double *field_params; /* HOST pointer, n_field_params doubles (Python-float args) */
  int32_t n_field_params;
  const void *field_weights;  /* DEVICE (or host for *_host) pointer to large parameters (MLP weights), dtype-typed */
  int64_t n_field_weights;

  /* batch and integration region; per-trajectory t0/t1 arrays are optional
   * ("vmappable everything, including the region of integration", README.md:10) */
  int64_t n_traj;
  const void *y0;             /* [N, d] */
  double t0, t1;
  const void *t0_per_traj;    /* [N] or NULL */
  const void *t1_per_traj;    /* [N] or NULL */
  double dt0;                 /* NaN == None (pid.py:328-340 -> 0.01 through diffeqsolve) */

  /* stepsize_controller */
  int32_t controller;         /* dfx_controller */
  double rtol, atol, pcoeff, icoeff, dcoeff, safety, factormin, factormax; /* pid.py:299-310 */
  double dtmin, dtmax;        /* NaN == None */
  int32_t force_dtmin;
  double error_order;         /* NaN == solver.error_order(terms) (base.py:97-120) */
  int32_t hairer_initial_step; /* dt0 == None only.  0: first trial step 0.01, which is what diffeqsolve does today
                                * (pid.py:48-49 is reached with WrapTerm-wrapped terms);
                                * 1: the Hairer II.4 starting step coded at pid.py:51-81 (ODE solves only) */

  /* ClipStepSizeController(controller, step_ts, jump_ts) (_step_size_controller/clip.py:120-428; also
   * PIDController(step_ts=, jump_ts=), pid.py:88-97): times that must be stepped to exactly / stepped around.
   * Sorted ascending, user time, time dtype; shared by all trajectories; NULL when unused. */
  const void *step_ts;
  int32_t n_step_ts;
  const void *jump_ts;
  int32_t n_jump_ts;
  int32_t store_rejected_steps;    /* 0 = None; else the length of the rejected-times stack: rejected step ends are
                                    * revisited by later steps (clip.py:292-299, 398-424) */

  /* saveat (_saveat.py:22-26, 72-76) and max_steps (_integrate.py:904) */
  int32_t save_t0, save_t1, save_steps, save_dense;
  const void *save_ts;        /* [T] in the time dtype, or NULL */
  int32_t n_save_ts;
  int32_t max_steps;

  /* outputs; T_out = dfx_out_size(desc) (_integrate.py:1273-1293).  Unused slots are +inf
   * (_integrate.py:1296-1300 followed by ts * direction at 1480-1482). */
  void *ts_out;               /* [N, T_out] or NULL when T_out == 0 */
  void *ys_out;               /* [N, T_out, d] */
  int32_t *stats;             /* [N, 3] num_steps, num_accepted_steps, num_rejected_steps (1518-1524) */
  int32_t *result;            /* [N] dfx_result */
  int32_t *save_count;        /* [N] filled output slots, or NULL */
  /* SaveAt(dense=True): DenseInterpolation(ts, infos) (_integrate.py:1315-1323, 529-540) */
  void *dense_ts;             /* [N, max_steps+1] */
  void *dense_y0;             /* [N, max_steps, d] */
  void *dense_y1;             /* [N, max_steps, d] */
  void *dense_k;              /* [N, max_steps, s, d]; NULL for two-point interpolants (Euler, ShARK) */
  int32_t *dense_count;       /* [N] accepted steps stored (ts_size - 1) */
  /* final state (Solution.ys[-1] for SaveAt(t1=True) duplicates this; kept for multi-GPU gathers) */
  void *y_final;              /* [N, d] or NULL */
  void *t_final;              /* [N] or NULL */

  /* VirtualBrownianTree(t0, t1, tol, shape=(), key, levy_area) per trajectory (tree.py:245-301) */
  int32_t levy_area;          /* dfx_levy; DFX_LEVY_NONE for ODEs */
  const uint32_t *bm_keys;    /* [N, 2] the keys the user passed to VirtualBrownianTree */
  double bm_t0, bm_t1, bm_tol;
  int32_t threefry_partitionable; /* jax_threefry_partitionable (default True since JAX 0.5) */
  int32_t bm_dim;                 /* 0: VirtualBrownianTree(shape=()); m > 0: shape=(m,).  Either way the tree has ONE leaf
                                   * (a tuple of ints is a single ShapeDtypeStruct, tree.py:291-295) keyed split(key, 1)[0]
                                   * (tree.py:301, _misc.py:128-133); each node draws jr.normal(key, (m,)), so the m
                                   * components share the key path.  Kernels: the OU functor with dim 2, 3 (diagonal
                                   * diffusion, m == dim). */

  /* Event(cond_fn, root_finder, direction): _event.py:13-118, _integrate.py:542-633 (detection), 691-821 (root find, unsave).
   * Up to DFX_MAX_EVENTS registered condition functions (the flattened PyTree `cond_fn`; the first one that triggers on a
   * step wins, _integrate.py:619-626), evaluated in the solver's (direction-normalised) time like the reference's
   * cond_fn(tprev, y, ...):
   *   DFX_EVENT_AFFINE        c(t, y) = w . y + wt * t + b        params [w[0..d), b, wt]  real-valued: sign change
   *   DFX_EVENT_STEADY_STATE  rms(f(t, y)) < atol + rtol * rms(y)  params [rtol, atol]      boolean (_event.py:120-170)
   *   DFX_EVENT_USER          the functor's own event<R>(params, j, t, y) params [j]               real-valued: sign change
   *                           (user functors with kUserEvents > 0, e.g. fields.CudaField(events=[...]); the root find
   *                           differentiates it along the interpolant by central differences)
   * event_params holds the parameters of event 0, 1, ... back to back (host doubles).
   * event_direction[i]: 0 = None (any crossing), 1 = True (upcrossing), 2 = False (downcrossing).
   * event_root_find: 0 = root_finder None (the solve ends at the end of the triggering step); 1 = Newton(event_rtol,
   *   event_atol) on the step's local interpolant, bracketed to the step, saved values after the event time removed. */
  int32_t n_events;                                 /* 0 = no event */
  int32_t event_kind[4], event_direction[4];        /* DFX_MAX_EVENTS == 4 */
  int32_t event_root_find;
  const double *event_params; int32_t n_event_params; /* host */
  double event_rtol, event_atol;

  /* Resuming a solve: diffeqsolve(..., solver_state=, controller_state=, made_jump=) and SaveAt(solver_state=True,
   * controller_state=True, made_jump=True) (_integrate.py:1250-1271, 1489-1500).  One record per trajectory, [N, 5 + d]:
   *   [0] prev_inv_scaled_error  [1] prev_prev_inv_scaled_error  [2] at_dtmin   (PIDController state, pid.py:388-392)
   *   [3] made_jump   [4] first_step   [5 ..] the carried FSAL derivative       (solver state, runge_kutta.py:415-444)
   * state_in_flags: bit 0 controller_state passed, bit 1 solver_state passed, bit 2 made_jump passed. */
  const void *state_in; int32_t state_in_flags;
  void *state_out;

  /* dfx_ensemble_solve_host only: optional DEVICE buffers (on the device the call runs on) that receive
   * the final states / times ([N, d] / [N]) in addition to the host outputs, so that a collective can follow the call without a
   * second H2D (the multi-GPU gather of SURVEY section 8e: NCCL all_gather of the finals).  Ignored by the device call,
   * where y_final / t_final are device pointers already. */
  void *y_final_device;
  void *t_final_device;

  /* Ensemble totals, reduced inside the solve kernel (the statistics half of the multi-GPU gather): [4] int64 =
   * sum of num_steps, sum of num_accepted_steps, number of trajectories whose result is neither successful nor event_occurred
   * (i.e. not is_okay, _solution.py:52-62), max num_steps of one
   * trajectory.  `totals` lives where the other outputs live (device for dfx_ensemble_solve, host for *_host);
   * `totals_device` is the optional device copy for the host call, like y_final_device.  NULL = not wanted. */
  int64_t *totals;
  int64_t *totals_device;

  /* SaveAt(dense=True): 0 = every unfilled slot of dense_ts / dense_y0 / dense_y1 / dense_k is written +inf by the solve
   * (the reference's buffers, _integrate.py:1296-1300, 1320-1322); 1 = the tails are left unwritten (dense_count says how
   * many records are valid; dfx_dense_evaluate / _derivative never read beyond them) and dfx_dense_pad() fills them on
   * demand.  At BASELINE config 3 the padding is 63 % of the bytes the eager layout writes. */
  int32_t dense_lazy_padding;

  /* Fused gather of the final states over NVLink peer memory (multi-GPU, one process per GPU).  When
   * n_peers > 0 the solve kernel stores every trajectory's final state / time, the moment it is finalised, into row
   * (peer_row_offset + i) of EACH of the n_peers buffers below - device pointers valid on this device: this rank's own
   * global buffer and its peers' buffers opened with dfx_peer_open (CUDA IPC; P2P stores travel over NVLink / NVSwitch).
   * The all_gather of the finals thereby overlaps the solve; what remains after the kernel is a barrier.  [N_total, d] /
   * [N_total] in the state dtype. */
  int32_t n_peers;                       /* 0 = off; at most DFX_MAX_PEERS */
  int64_t peer_row_offset;               /* first global row of this rank's block */
  void *peer_y_final[8];                 /* DFX_MAX_PEERS == 8 */
  void *peer_t_final[8];

  /* Per-trajectory `args` (the vmapped `args` of diffeqsolve(..., args=...), _integrate.py:896: a parameter sweep across the
   * ensemble): [n_traj, n_traj_args] dtype-typed values, DEVICE (or host for *_host) pointer; trajectory i evaluates the
   * functor with parameters traj_args[i, :] instead of field_params.  Only for functors compiled with kPerTrajArgs
   * (fields.CudaField registers such a variant under its own field id); NULL = off. */
  const void *traj_args;
  int32_t n_traj_args;
} dfx_solve_desc;

/* ---- library ---- */
int dfx_abi_version(void);
const char *dfx_last_error(void);          /* thread-local message for the last negative return */
int dfx_device_count(void);                /* 0 when no CUDA device / driver */

/* ---- registry introspection ---- */
int dfx_num_stages(int solver_id);         /* ButcherTableau.num_stages (runge_kutta.py:164) */
int dfx_solver_order(int solver_id);       /* solver.order(terms) */
int dfx_field_dim(int field_id);           /* fixed dim of a registered functor, 0 = any */
int dfx_has_kernel(int field_id, int dim, int solver_id, int dtype, int levy_area);

/* replaces _allocate_output (_integrate.py:1273-1293): number of output slots T_out */
int dfx_out_size(const dfx_solve_desc *desc);

/* replaces jax.vmap(diffeqsolve) forward pass (_integrate.py:888-1543 incl. loop 302-885,
 * runge_kutta.py:446-1203, srk.py:335-671, pid.py:316-567, constant.py:30-104, interpolants,
 * tree.py:326-773).  Device pointers; enqueues on `cuda_stream`. */
int dfx_ensemble_solve(const dfx_solve_desc *desc, void *cuda_stream);

/* same call with HOST buffers: copies inputs H2D, solves, copies outputs D2H, synchronises.
 * `device` selects the CUDA device.  This is the end-to-end (`e2e`) path of bench.py.
 * Pass pinned buffers for full PCIe speed.  Large adaptive SaveAt(t1=True) solves run as one launch
 * with the transfers chunked alongside it; everything else as up to 8 pipelined launches. */
int dfx_ensemble_solve_host(const dfx_solve_desc *desc, int device);

/* replaces VirtualBrownianTree.evaluate(t0, t1, use_levy=True) vmapped over keys
 * (tree.py:326-354).  ta/tb: [n] if per_traj_times else [1].  bm_dim: 0 = shape (), m = shape (m,) (m <= 8).
 * W, H: [n] or [n, m]; H may be NULL. */
int dfx_vbt_evaluate(int dtype, int levy_area, int partitionable, int64_t n, const uint32_t *keys,
                     double bm_t0, double bm_t1, double bm_tol, const void *ta, const void *tb,
                     int per_traj_times, void *W, void *H, int bm_dim, void *cuda_stream);

/* jax.random primitives the tree relies on, exposed for known-answer tests
 * (jax/_src/prng.py threefry2x32 / split / random_bits+normal). */
int dfx_threefry2x32(int64_t n, const uint32_t *keys /*[n,2]*/, const uint32_t *ctrs /*[n,2]*/,
                     uint32_t *out /*[n,2]*/, void *cuda_stream);
int dfx_random_split(int64_t n, const uint32_t *keys /*[n,2]*/, int num, int partitionable,
                     uint32_t *out /*[n,num,2]*/, void *cuda_stream);
/* jax.random.normal(key, shape, dtype) per key: m = 0 -> shape (), out [n]; m > 0 -> shape (m,), out [n, m] (m <= 8).
 * The float side (log1p, erf_inv) is one explicitly sequenced IEEE evaluation (csrc/prng.cuh). */
int dfx_random_normal(int dtype, int64_t n, const uint32_t *keys /*[n,2]*/, int partitionable,
                      void *out, int m, void *cuda_stream);

/* replaces DenseInterpolation.evaluate vmapped (_global_interpolation.py:335-355):
 * trajectory i is evaluated at tq[i, 0..nq).  out: [N, nq, d]. */
int dfx_dense_evaluate(int dtype, int solver_id, int64_t n_traj, int dim, int max_steps,
                       const void *dense_ts, const void *dense_y0, const void *dense_y1,
                       const void *dense_k, const int32_t *dense_count, double direction,
                       const void *tq, int nq, void *out, void *cuda_stream);
/* replaces DenseInterpolation.derivative vmapped (_global_interpolation.py:357-368; the local interpolants' derivative is
 * the jax.jvp tangent of their evaluate, _path.py): same arguments, out[i, q, :] = d/dt of the interpolant at tq[i, q]. */
int dfx_dense_derivative(int dtype, int solver_id, int64_t n_traj, int dim, int max_steps,
                         const void *dense_ts, const void *dense_y0, const void *dense_y1,
                         const void *dense_k, const int32_t *dense_count, double direction,
                         const void *tq, int nq, void *out, void *cuda_stream);

/* dst_device[0..n) = *src_device on `cuda_stream`: a device-resident scalar (an unbatched traced t0 / t1 reaching the
 * jax.ffi handler) turned into the per-trajectory array the descriptor takes. */
int dfx_broadcast_device_scalar(int dtype, int64_t n, const void *src_device, void *dst_device, void *cuda_stream);

/* fills the unfilled tails of dense buffers produced with dense_lazy_padding = 1 with +inf (device pointers) */
int dfx_dense_pad(int dtype, int solver_id, int64_t n_traj, int dim, int max_steps, void *dense_ts, void *dense_y0,
                  void *dense_y1, void *dense_k, const int32_t *dense_count, void *cuda_stream);

/* Peer memory for the fused gather (dfx_solve_desc.peer_*): a device allocation that other processes of the same node can
 * map.  dfx_peer_alloc: cudaMalloc on the current device + its 64-byte CUDA IPC handle (send it to the peers by any
 * means, e.g. torch.distributed.all_gather_object); dfx_peer_open: map a peer's allocation into this process (peer access
 * enabled lazily); dfx_peer_close / dfx_peer_free undo them. */
int dfx_peer_alloc(int64_t bytes, void **device_ptr, void *ipc_handle_64_bytes);
int dfx_peer_open(const void *ipc_handle_64_bytes, void **device_ptr);
int dfx_peer_close(void *device_ptr);
int dfx_peer_free(void *device_ptr);

/* measured FMA-pipe peaks for the roofline denominators (dependent-free FMA chains);
 * returns TFLOP/s (2 flop per FMA) or a negative dfx_error. */
double dfx_measure_fma_peak(int dtype, int device);
/* INT32 ALU peak (threefry-like add/rot/xor chain), Tera-ops/s */
double dfx_measure_int_peak(int device);

/* number of kernels this library launched on behalf of the calling thread since the last
 * reset (bench.py's `gpu_launches`). */
int64_t dfx_launch_count(void);
void dfx_reset_launch_count(void);

/* ---- user functor registration (see diffrax_b200/csrc/register_field.cuh) ---- */
typedef int (*dfx_launcher_fn)(const dfx_solve_desc *desc, void *cuda_stream);
int dfx_register_launcher(int field_id, int dim, int solver_id, int dtype, int levy_area,
                          dfx_launcher_fn fn);

#ifdef __cplusplus
}
#endif
#endif /* DIFFRAX_B200_H */
