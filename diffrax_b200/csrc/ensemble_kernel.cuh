// ensemble_kernel.cuh - the persistent adaptive ensemble integrator kernel (K1/K2/K4 of SURVEY.md §2.1).
//
// One kernel replaces, per trajectory, the whole body of the reference's vmapped diffeqsolve:
//   _integrate.py:302-885   loop / body_fun_aux: step -> adapt -> clip to t1 -> keep/restore -> counters -> save
//   runge_kutta.py:446-1203 explicit single-tableau ERK step with FSAL / SSAL
//   euler.py:46-59, srk.py:335-671 (+ shark.py:10-30) additive-noise SRK step
//   pid.py:316-392, 394-567 PID controller; constant.py:30-104 ConstantStepSize
//   _local_interpolation.py / tsit5.py / dopri5.py / dopri8.py interpolants for SaveAt(ts)
//   _brownian/tree.py       VirtualBrownianTree increments (vbt.cuh)
//   _solver/base.py:250-346 HalfSolver (HalfOf<Inner>); and, in the EXTRA instantiation only:
//   clip.py:120-428 ClipStepSizeController, pid.py:33-81 Hairer starting step, _event.py + _integrate.py:542-633, 691-821
//   Events with a Newton root find on the local interpolant, _integrate.py:1250-1271 resumed solver / controller state
//
// Mapping to the hardware
//   * one trajectory per thread; y, the s stage values k[s][d], the FSAL derivative and the
//     controller state stay in registers for the whole solve (stage loops fully unrolled, Butcher
//     coefficients read from __constant__ memory as immediate c[bank][offset] operands);
//   * HBM is touched only to read y0 once and to write saved outputs;
//   * the adaptive loops diverge per trajectory, so the grid is persistent (a multiple of the SM
//     count) and a lane that finishes its trajectory claims the next one from a global work queue
//     (warp-aggregated atomicAdd: one atomic per refilling warp; finished lanes wait until
//     `refill_batch` of them share one finalize + refill pass).  The still active lanes therefore
//     stay compacted in full warps until the queue drains; results do not depend on the
//     lane <-> trajectory assignment because trajectories are independent.
//   * accept / reject is a select, not a branch: every lane executes the same instruction stream.
#pragma once
#include <type_traits>
#include "common.cuh"
#include "fields.cuh"
#include "interp.cuh"
#include "tableaux.cuh"
#include "vbt.cuh"

namespace dfx {

struct EulerSolver {
  static constexpr int kId = DFX_EULER;
  static constexpr int S = 1;
  static constexpr int kOrder = 1;
  static constexpr bool kFsal = false, kSsal = false;
  static constexpr int kInterp = kInterpLinear;
  static constexpr bool kIsTableau = false;
};
struct SharkSolver {
  static constexpr int kId = DFX_SHARK;
  static constexpr int S = 2;
  static constexpr int kOrder = 2;
  static constexpr bool kFsal = false, kSsal = false;
  static constexpr int kInterp = kInterpLinear;
  static constexpr bool kIsTableau = false;
};
// HalfSolver(inner), _solver/base.py:250-346: every step is two half steps of `inner` (the result) and one full step (the
// dense_info and, through |y1 - y1_alt|, the error estimate).  Inherits the inner solver's tableau / interpolant.
template <class Inner>
struct HalfOf : Inner {
  static constexpr int kId = DFX_HALF_SOLVER | Inner::kId;
  static constexpr int kInnerId = Inner::kId;
  static constexpr int kOrder = Inner::kOrder + 1;  // error order on an ODE (base.py:296-299)
  static constexpr bool kHalf = true;
};
// number of independent Brownian components a field is driven by (Field::kNoise when it declares one, else 1)
template <class F, class = void> struct NoiseDim { static constexpr int value = 1; };
template <class F> struct NoiseDim<F, std::void_t<decltype(F::kNoise)>> { static constexpr int value = F::kNoise; };
// matrix-valued diffusion: Field::noise_prod(fp, t, W[NW], c) = row c of g(t) . W instead of the scalar g(t) * W[c]
template <class F, class = void> struct MatrixNoise { static constexpr bool value = false; };
template <class F> struct MatrixNoise<F, std::void_t<decltype(F::kMatrixNoise)>> { static constexpr bool value = F::kMatrixNoise; };
// state-dependent (multiplicative) diffusion: Field::noise_prod(fp, t, y, W[NW], c) = component c of g(t, y) . W
template <class F, class = void> struct StateNoise { static constexpr bool value = false; };
template <class F> struct StateNoise<F, std::void_t<decltype(F::kStateNoise)>> { static constexpr bool value = F::kStateNoise; };
// condition functions of the functor's own (Event(cond_fn) with an arbitrary cond_fn): Field::event<R>(fp, j, t, y), j < kUserEvents
template <class F, class = void> struct UserEvents { static constexpr int value = 0; };
template <class F> struct UserEvents<F, std::void_t<decltype(F::kUserEvents)>> { static constexpr int value = F::kUserEvents; };
// per-trajectory functor parameters (the vmapped `args`): Field::P<R> is { R p[kNumParams]; } and trajectory i uses traj_args[i, :]
template <class F, class = void> struct PerTrajArgs { static constexpr bool value = false; };
template <class F> struct PerTrajArgs<F, std::void_t<decltype(F::kPerTrajArgs)>> { static constexpr bool value = F::kPerTrajArgs; };
template <class T, class = void> struct IsHalf { static constexpr bool value = false; };
template <class I> struct IsHalf<HalfOf<I>> { static constexpr bool value = true; };
template <class T> struct InnerId { static constexpr int value = T::kId; };
template <class I> struct InnerId<HalfOf<I>> { static constexpr int value = I::kId; };
template <class T, class = void> struct IsTableau { static constexpr bool value = true; };
template <> struct IsTableau<EulerSolver> { static constexpr bool value = false; };
template <> struct IsTableau<SharkSolver> { static constexpr bool value = false; };
template <class I> struct IsTableau<HalfOf<I>> { static constexpr bool value = IsTableau<I>::value; };

template <class R>
struct SolveParams {
  long long n_traj;
  const R *y0;
  const R *t0_arr, *t1_arr;
  R t0, t1;
  R dt0;
  int has_dt0;
  int controller;
  R rtol, atol, safety, factormin, factormax, dtmin, dtmax;
  int has_dtmin, has_dtmax, force_dtmin;
  R coeff1, coeff2, coeff3;  // PID exponents (pid.py:512-514), computed in double on the host
  int use_c1, use_c2, use_c3;
  const R *step_ts, *jump_ts;  // ClipStepSizeController (clip.py): sorted, user time; EXTRA instantiation only
  int n_step_ts, n_jump_ts;
  R *reject_ts; int n_reject;  // store_rejected_steps: per-trajectory stack of rejected step ends, [N, n_reject] scratch (EXTRA only)
  int hairer;    // dt0 == None: use the Hairer starting step of pid.py:51-81 instead of the constant 0.01
  R inv_error_order;
  int fast_pid;  // pcoeff == dcoeff == 0, icoeff == 1 and error_order == solver order: pure I-controller fast path
  int save_t0, save_t1, save_steps, save_dense;
  const R *save_ts;
  int n_save_ts, max_steps, out_size;
  R *ts_out, *ys_out;
  int *stats, *result, *save_count;
  R *dense_ts, *dense_y0, *dense_y1, *dense_k;
  int *dense_count;
  // Event (EXTRA instantiation only): kind, direction (0 any / 1 up / 2 down), Newton root find on the local interpolant
  int n_events, event_kind[DFX_MAX_EVENTS], event_dir[DFX_MAX_EVENTS], event_root;
  R ev_w[DFX_MAX_EVENTS][4], ev_b[DFX_MAX_EVENTS], ev_wt[DFX_MAX_EVENTS];  // affine: w . y + wt t + b
  R ev_ss_rtol[DFX_MAX_EVENTS], ev_ss_atol[DFX_MAX_EVENTS];               // steady state
  int ev_user[DFX_MAX_EVENTS];                                            // user: index of the functor's condition
  R ev_rtol, ev_atol;                                                     // Newton root finder
  const R *state_in; R *state_out; int state_in_flags;  // resumed / returned controller + solver state, [N, 5 + d] (EXTRA only)
  int refill_batch;  // finished lanes wait until this many can be finalised + refilled in one pass (>= 1)
  int dense_smem_offset;  // bytes of dynamic shared memory in front of the dense staging records (the VBT descent cache)
  int pad_vec, flush_vec;  // 32-byte stores for the +inf padding / the dense record flush
  int dense_cs;      // dense records with st.global.cs (evict-first) instead of write-back stores
  int dense_coop;    // SaveAt(dense): stage records through shared memory and store them warp-cooperatively (launcher provides the smem)
  int dense_lazy;    // SaveAt(dense): leave the unfilled tails unwritten (dfx_dense_pad fills them on demand)
  int dense_vec_ok;  // dense_y0 / dense_y1 / dense_k base pointers are 32-byte aligned (rows then are, when their size allows)
  R *y_final, *t_final;
  // fused gather over peer memory (dfx_solve_desc.peer_*): final states / times also go to row peer_row0 + i of every peer buffer
  int n_peers; long long peer_row0; R *peer_y[DFX_MAX_PEERS]; R *peer_t[DFX_MAX_PEERS];
  long long *totals;  // [4] or null: sums of attempted / accepted steps, failed trajectories, max steps of one trajectory (zeroed by the launcher)
  unsigned long long *work_counter;
  // Host-pipelined mode (dfx_ensemble_solve_host; SaveAt(t1=True) instantiation only): ONE launch over the whole batch
  // while the copy engines bring the inputs in and take the results out in chunks of pipe_chunk_len trajectories.
  const unsigned *pipe_in_ready;  // device words [in_ready, abort]: number of input chunks resident so far (bumped by a 4-byte H2D copy
                                  // that is stream-ordered after the chunk's data)
  unsigned *pipe_done;            // device: finalised trajectories per chunk
  unsigned *pipe_host_flags;      // mapped pinned host memory: word c becomes 1 when every result of chunk c is written
  int pipe_chunk_len;             // multiple of 32 trajectories, so no 128-byte line straddles two chunks
  const uint32_t *keys;
  VbtParams vbt;
  const R *traj_args;  // [n_traj, n_traj_args] per-trajectory functor parameters, or null
  int n_traj_args;
};

#ifndef DFX_BLOCK_THREADS
#define DFX_BLOCK_THREADS 128
#endif
constexpr int kBlockThreads = DFX_BLOCK_THREADS;

// Occupancy target handed to ptxas.  The stage values k[S][D] must stay in registers; the rest of the state needs
// ~64 more (+ the interpolation / save bookkeeping of the RICH variant, + the Brownian tree for SDEs).  Measured on
// C2 (Lorenz/Dopri5/fp64): 3 CTAs/SM (140 regs) 4.53 ms, 6 CTAs/SM (79 regs, no spills) 3.71 ms, 8 CTAs/SM
// (64 regs, spills) 3.87 ms - so the default asks for as many CTAs as comfortably fit (capped at 6) and
// MinBlocksOverride pins the measured optimum for the benchmark configurations.
template <class R, class Field, class Solver, int LEVY, bool RICH>
struct MinBlocksOverride { static constexpr int value = 0; };
template <> struct MinBlocksOverride<double, LorenzField, Dopri5, 0, false> { static constexpr int value = 6; };
// the per-thread MLP evaluation keeps the 128 hidden activations in registers
template <class R, int D, int W, class Solver, int LEVY, bool RICH>
struct MinBlocksOverride<R, MlpField<D, W>, Solver, LEVY, RICH> { static constexpr int value = 2; };

template <class R, class Field, class Solver, int LEVY, bool RICH>
constexpr int min_blocks_per_sm() {
#ifdef DFX_MIN_BLOCKS
  return DFX_MIN_BLOCKS;
#else
  if (MinBlocksOverride<R, Field, Solver, LEVY, RICH>::value > 0) return MinBlocksOverride<R, Field, Solver, LEVY, RICH>::value;
  const int words = (int)sizeof(R) / 4;
  const int budget = Solver::S * Field::kDim * words + 64 + (RICH ? 32 : 0) + (LEVY != 0 ? 24 * words : 0);
  const int blocks = 65536 / (kBlockThreads * budget);
  return blocks < 1 ? 1 : (blocks > 6 ? 6 : blocks);
#endif
}

// +inf into row[start, len), one warp, coalesced streaming stores
template <class R> __device__ __forceinline__ void pad_tail(R *row, long long start, long long len, int lane, bool vec = true) {
  if (!vec) {
    for (long long i = start + lane; i < len; i += 32) st_cs(&row[i], Num<R>::inf());
    return;
  }
  constexpr int VW = 32 / (int)sizeof(R);  // elements per 32-byte store
  R *q = row + start;
  long long n = len - start;
  if (n <= 0) return;
  long long head = (long long)(((32u - (unsigned)((uintptr_t)q & 31u)) & 31u) / sizeof(R));  // up to the 32-byte boundary
  if (head > n) head = n;
  if (lane < head) st_cs(q + lane, Num<R>::inf());
  q += head;
  n -= head;
  const long long nvec = n / VW;
  for (long long i = lane; i < nvec; i += 32) st32B_fill_cs(q + i * VW, Num<R>::inf());  // 1 KB per warp instruction
  const long long done = nvec * VW;
  if (done + lane < n) st_cs(q + done + lane, Num<R>::inf());
}

// Claim the next trajectory for every lane of the warp that needs one: one atomic per warp.
__device__ __forceinline__ long long claim_work(bool need, unsigned long long *counter, int lane) {
  const unsigned m = __ballot_sync(kFullMask, need);
  if (m == 0) return -1;
  const int leader = __ffs(m) - 1;
  unsigned long long base = 0;
  if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
  base = __shfl_sync(kFullMask, base, leader);
  return need ? (long long)(base + __popc(m & ((1u << lane) - 1u))) : -1;
}

// ---- ClipStepSizeController helpers (clip.py:53-97).  After wrap(direction) the controller sees sort(ts * direction):
// element i is direction > 0 ? ts[i] : -ts[n-1-i].
template <class R> __device__ __forceinline__ R clip_at(const R *ts, int n, int i, R direction) {
  return direction > R(0) ? ts[i] : -ts[n - 1 - i];
}
template <class R> __device__ __forceinline__ R clip_get_t(const R *ts, int n, int i, R direction) {  // _get_t
  return (n == 0 || i >= n) ? Num<R>::inf() : clip_at(ts, n, i, direction);
}
template <class R> __device__ __forceinline__ R next_after_up(R x) {  // eqxi.nextafter(x) == nextafter(x, +inf), finite x
  return -prev_n<R>(-x, 1);
}
template <class R> __device__ __forceinline__ bool clip_contains(const R *ts, int n, R x, R direction) {  // jnp.any(x == ts)
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (clip_at(ts, n, mid, direction) < x) lo = mid + 1; else hi = mid; }
  return lo < n && clip_at(ts, n, lo, direction) == x;
}
template <class R> __device__ __forceinline__ R clip_bump(R next_t0, const R *ts, int n, R direction, bool &made_jump) {  // _bump_next_t0
  const R nn = next_after_up(next_t0);
  const bool mj1 = clip_contains(ts, n, nn, direction), mj2 = clip_contains(ts, n, next_t0, direction);
  if (mj1) next_t0 = next_after_up(nn);
  if (mj2) next_t0 = nn;
  made_jump = mj1 || mj2;
  return next_t0;
}
template <class R> __device__ __forceinline__ int clip_find_idx(R t, const R *ts, int n, int hint, R direction) {  // _find_idx_with_hint
  int i = hint;
  while (i < n && clip_at(ts, n, i, direction) <= t) ++i;
  while (i > 0 && clip_at(ts, n, i - 1, direction) > t) --i;
  return i;
}

// Template parameters
//   R      working dtype (state == time dtype)
//   Field  vector-field functor (fields.cuh)
//   Solver generated tableau struct (tableaux.cuh), EulerSolver or SharkSolver
//   LEVY   dfx_levy: 0 ODE, 1 BrownianIncrement, 2 SpaceTimeLevyArea
//   RICH   false: SaveAt(t1=True) only (the C2/C4/C5 fast path); true: every SaveAt mode
//   EXTRA  (RICH only) the rarely used machinery: ClipStepSizeController, the Hairer starting step, Events.  Kept out of
//          the plain SaveAt kernels because it costs them registers and instruction-cache footprint (Dopri8 dense: 216 ->
//          255 registers with spills when it was compiled in)
//   SPEC   (fp64 ODE SaveAt(t1) solves only) the controller configuration is known at compile time: PIDController with the
//          default pure I-controller at the solver's order (the fast path) and no dtmin / dtmax.  Removes the uniform
//          branches and parameter loads of the general controller from the step loop (C2: -2 % time).
//   MOREB  extra CTAs per SM asked of ptxas on top of the default occupancy target (SPEC SaveAt(t1) kernels only): a batch
//          that is a little larger than the default grid's lanes runs as ONE resident generation at the higher occupancy
//          instead of one generation plus a mostly empty second one (the sharded 2^20 / 8 GPUs case: 131 072 trajectories
//          vs 113 664 lanes at 6 CTAs/SM, 132 608 at 7)
template <class R, class Field, class Solver, int LEVY, bool RICH, bool EXTRA = false, bool SPEC = false, int MOREB = 0>
__global__ void __launch_bounds__(kBlockThreads, min_blocks_per_sm<R, Field, Solver, LEVY, RICH>() + MOREB)
ensemble_kernel(const SolveParams<R> p, const typename Field::template P<R> fp_in) {
  // the functor's parameters: the launch-wide kernel argument, or (functors compiled with kPerTrajArgs) a per-lane copy that is
  // reloaded from traj_args[idx, :] whenever the lane claims a trajectory; the copy is dead code for every other functor
  [[maybe_unused]] typename Field::template P<R> fp_lane = fp_in;
  const typename Field::template P<R> &fp = PerTrajArgs<Field>::value ? fp_lane : fp_in;
  constexpr int D = Field::kDim;
  constexpr int S = Solver::S;
  constexpr bool SDE = LEVY != DFX_LEVY_NONE;
  constexpr bool TAB = IsTableau<Solver>::value;
  constexpr bool FSAL = Solver::kFsal && !SDE;
#ifndef DFX_OPT_CHAIN_Y0
#define DFX_OPT_CHAIN_Y0 1
#endif
#ifndef DFX_OPT_ABSMAX_FP64
#define DFX_OPT_ABSMAX_FP64 1   // max(|y0|, |y1|) of the error scale on the FP64 pipe (DSETP + select) instead of 7 integer-ALU ops:
                                // the issue port, not the FP64 pipe, is the tighter limit of the step loop (C2: -1.8 % time)
#endif
#ifndef DFX_OPT_EARLY_STAGE
#define DFX_OPT_EARLY_STAGE 0   // SaveAt(dense): write each stage value to the staging record as soon as it exists
#endif
#ifndef DFX_OPT_FAST_PID
#define DFX_OPT_FAST_PID 1      // the division- / pow-free I-controller path of fp64 ODE solves (and with it the SPEC instantiations)
#endif
#ifndef DFX_OPT_LAST_STAGE_F
#define DFX_OPT_LAST_STAGE_F 1
#endif
  // Instruction-count reductions of the ODE step (see the stage loop).  The chained form is limited to pairs with <= 7
  // stages: the 14-stage Dopri8 sums are kept in the reference's "sum, then add y0" order (rtol 1e-12 territory).
  [[maybe_unused]] constexpr bool kChainY0 = DFX_OPT_CHAIN_Y0 && TAB && !SDE && Solver::S <= 7;
  [[maybe_unused]] constexpr bool kLastStageF = DFX_OPT_LAST_STAGE_F && TAB && FSAL && Solver::kSsal;
  constexpr int INTERP = Solver::kInterp;
  constexpr int NW = NoiseDim<Field>::value;  // Brownian components: 1 (shape=()) or D (shape=(D,), diagonal diffusion)
  constexpr bool DENSE_K = INTERP != kInterpLinear;
  constexpr bool FAST_PID = DFX_OPT_FAST_PID && TAB && !SDE && sizeof(R) == 8;  // fp64 ODE solves: division-/pow-free I-controller path

  // ---- per-lane trajectory state (registers) ----
  bool active = false, exhausted = false;
  [[maybe_unused]] bool pipe_aborted = false;  // host-pipelined mode: the host gave up delivering inputs
  int idx_i = -1;  // trajectory index (n_traj < 2^31 is checked on the host); widened at each use
  R y[D], f_fsal[D];
  R tprev = R(0), tnext = R(0), t0 = R(0), t1 = R(0), direction = R(1), t1_clip_floor = R(0);
  R pid_inv = R(1), pid_prev_inv = R(1);
  bool at_dtmin = false;
  int cs_num_steps = 0;  // ConstantStepSize: steps_completed (constant.py:84) is num_steps + 1, every step being accepted
  int num_steps = 0, num_accepted = 0, result = DFX_RESULT_SUCCESSFUL;
  int save_index = 0, saveat_ts_index = 0, dense_index = 0;
  [[maybe_unused]] int step_index = 0, jump_index = 0;  // ClipStepSizeController state (EXTRA only)
  [[maybe_unused]] bool made_jump = false;
  [[maybe_unused]] int reject_index = 0;  // ClipStepSizeController(store_rejected_steps=K): top of the stack, K = empty
  [[maybe_unused]] R event_value[DFX_MAX_EVENTS] = {};  // Event: the cond_fns at the previous state (EXTRA only)
  BrownianTree<R, LEVY == DFX_LEVY_SPACE_TIME, NW> bm;  // one tree; shape (m,) = m components sharing the key path
#pragma unroll
  for (int c = 0; c < D; ++c) { y[c] = R(0); f_fsal[c] = R(0); }

  const R sqrt_d = (R)sqrt((double)D);

  // Event condition (see dfx_solve_desc): affine  w . y + wt t + b,  or steady state  rms(f) < atol + rtol rms(y)  (1 = True)
  [[maybe_unused]] auto event_cond = [&](int i, R t, const R (&yy)[D], R dir) -> R {
    if (p.event_kind[i] == DFX_EVENT_AFFINE) {
      R v = R(0);
#pragma unroll
      for (int c = 0; c < D; ++c) v += p.ev_w[i][c < 4 ? c : 3] * yy[c];
      return v + p.ev_wt[i] * t + p.ev_b[i];
    }
    if (p.event_kind[i] == DFX_EVENT_USER) {
      if constexpr (UserEvents<Field>::value > 0) return Field::template event<R>(fp, p.ev_user[i], t, yy);
      else return R(0);
    }
    R f[D], nf = R(0), ny = R(0);
    Field::template eval<R>(fp, t * dir, yy, f);
    if constexpr (D == 1) { nf = r_abs(f[0]); ny = r_abs(yy[0]); }
    else {
#pragma unroll
      for (int c = 0; c < D; ++c) { nf += f[c] * f[c]; ny += yy[c] * yy[c]; }
      nf = r_sqrt(nf) / sqrt_d; ny = r_sqrt(ny) / sqrt_d;
    }
    return (nf < p.ev_ss_atol[i] + p.ev_ss_rtol[i] * ny) ? R(1) : R(0);
  };

  // SaveAt(dense=True) staging: one record of kDenseRec values per lane, padded to an odd stride (conflict-free reads)
  constexpr int kDenseK = DENSE_K ? S * D : 0;
  constexpr int kDenseRec = kDenseK + 2 * D;
  // 32-byte chunk path: every piece of the record (k, y0, y1) is a whole number of 32-byte vectors, so a lane moves one
  // vector (2 x LDS.128 + 1 x STG.256) and a half / quarter warp moves a whole record
  constexpr int kVW = 32 / (int)sizeof(R);
  constexpr bool kDenseVec = (D % kVW == 0) && (kDenseK % kVW == 0) && (kDenseRec / kVW <= 16);
  constexpr int kDenseChunks = kDenseRec / kVW;                                      // vectors per record
  constexpr int kDenseGroup = kDenseChunks <= 4 ? 4 : (kDenseChunks <= 8 ? 8 : 16);  // lanes per record
  // staging stride per lane: odd (conflict-free 64-bit accesses) for the scalar path; for the chunk path a multiple of 16
  // bytes whose half is odd (records stay 16-byte aligned, per-lane element writes are 2-way conflicted at worst)
  constexpr int kDenseStride = kDenseVec ? (((kDenseRec * (int)sizeof(R) + 15) / 16) | 1) * 16 / (int)sizeof(R) : (kDenseRec | 1);
  // dynamic shared memory: [VBT descent cache (SDE kernels)] [dense staging records (RICH, SaveAt(dense))]
  extern __shared__ __align__(16) unsigned char dense_smem_raw[];
  [[maybe_unused]] R *dense_smem = reinterpret_cast<R *>(dense_smem_raw + p.dense_smem_offset);
  if constexpr (SDE) bm.attach_cache(reinterpret_cast<R *>(dense_smem_raw), p.vbt);

  __shared__ long long warp_totals[kBlockThreads / 32][4];  // attempted, accepted, failed, max steps (see p.totals)
  // lane id and this warp's totals slot as a 32-bit shared address, computed ONCE: left to the compiler, the special-register
  // reads (tid, shared window) behind them are re-issued at the top of every iteration of the step loop
  const int lane_id = threadIdx.x & 31;
  const uint32_t tot_addr = (uint32_t)__cvta_generic_to_shared(&warp_totals[threadIdx.x >> 5][0]);
  if (lane_id == 0) {
    asm volatile("st.shared.v2.u64 [%0], {%1, %1};" :: "r"(tot_addr), "l"(0ull) : "memory");
    asm volatile("st.shared.v2.u64 [%0], {%1, %1};" :: "r"(tot_addr + 16u), "l"(0ull) : "memory");
  }
  __syncwarp();

  for (;;) {
    // Finalising a trajectory and claiming + initialising the next one is ~300 instructions that the whole warp
    // issues for however few lanes need them; lanes finish at unrelated iterations, so doing it per lane costs about
    // 32 * 300 / (steps per trajectory * instructions per step) of the run (10 % on C2).  Finished lanes therefore
    // wait (idle) until `refill_batch` of them can share one pass, or nothing else is running.
    // a lane is `done` when its loop condition (_integrate.py:355-363, 685-687) has turned false
    // (SPEC: no dtmin, no events - `result` can only change when the trajectory is finalised)
    // `runnable` is evaluated once per iteration (here, and again only for lanes that are refilled below)
    bool runnable = (tprev < t1) && (num_steps < p.max_steps) && (SPEC || result == DFX_RESULT_SUCCESSFUL);
    const bool done = active && !runnable;
    const unsigned running = __ballot_sync(kFullMask, active && !done);
    const unsigned waiting = __ballot_sync(kFullMask, done);
    if (running == 0u || __popc(waiting) >= p.refill_batch) {
      // ---------------- finalize finished lanes ----------------
      if (done) {
        const long long idx = idx_i;
        if constexpr (RICH) {
          if (t0 == t1 && p.save_ts != nullptr) {  // _integrate.py:823-845
            for (int i = 0; i < p.n_save_ts; ++i) {
              const long long o = idx * (long long)p.out_size + save_index;
              p.ts_out[o] = t0 * direction;
  #pragma unroll
              for (int c = 0; c < D; ++c) p.ys_out[o * D + c] = y[c];
              save_index += 1;
            }
          }
        }
        {  // _integrate.py:847-877
          bool via_steps = false;
          if constexpr (RICH) {  // (the SaveAt(t1=True)-only instantiation has no steps and exactly one slot)
            if (p.save_steps == 1) via_steps = true;
            else if (p.save_steps > 1) via_steps = (num_accepted % p.save_steps) == 0;
          }
          // 862-872: with an event root finder the final value is (re)written whenever steps would have saved it
          const bool ev_rule = EXTRA && p.n_events != 0 && p.event_root;
          const bool pred = ev_rule ? (p.save_t1 || via_steps) : (p.save_t1 && !via_steps);
          if (pred && (!RICH || save_index < p.out_size)) {
            const long long o = idx * (long long)p.out_size + save_index;
            p.ts_out[o] = tprev * direction;
  #pragma unroll
            for (int c = 0; c < D; ++c) p.ys_out[o * D + c] = y[c];
            save_index += 1;
          }
        }
        if ((tprev < t1) && result == DFX_RESULT_SUCCESSFUL) result = DFX_RESULT_MAX_STEPS_REACHED;  // 883
        p.stats[idx * 3 + 0] = num_steps;
        p.stats[idx * 3 + 1] = num_accepted;
        p.stats[idx * 3 + 2] = num_steps - num_accepted;
        p.result[idx] = result;
        if (p.save_count) p.save_count[idx] = save_index;
        if (p.dense_count) p.dense_count[idx] = dense_index;
        if (p.y_final) {
  #pragma unroll
          for (int c = 0; c < D; ++c) p.y_final[idx * D + c] = y[c];
        }
        if (p.t_final) p.t_final[idx] = tprev * direction;
        if (p.n_peers != 0) {
          // the all_gather of the finals, fused: P2P stores over NVLink into every rank's global buffer (this rank's own
          // included), issued as trajectories finish, so the transfer rides along with the solve
          const long long g = p.peer_row0 + idx;
#pragma unroll 1
          for (int q = 0; q < p.n_peers; ++q) {
            R *yq = p.peer_y[q] + g * D;
#pragma unroll
            for (int c = 0; c < D; ++c) yq[c] = y[c];
            p.peer_t[q][g] = tprev * direction;
          }
        }
        if constexpr (!RICH) {
          if (p.pipe_done != nullptr) {  // release this trajectory's results; the last one of a chunk tells the host
            __threadfence();
            const int c = (int)(idx / p.pipe_chunk_len);
            const long long rest = p.n_traj - (long long)c * p.pipe_chunk_len;
            const unsigned cnt = (unsigned)(rest < p.pipe_chunk_len ? rest : (long long)p.pipe_chunk_len);
            // one atomic per group of lanes that finalise into the same chunk (fixed-step ensembles finish in waves)
            const unsigned peers = __match_any_sync(__activemask(), c);
            unsigned total = 0;
            if (lane_id == __ffs(peers) - 1) total = atomicAdd(p.pipe_done + c, (unsigned)__popc(peers)) + (unsigned)__popc(peers);
            if (total == cnt) {
              __threadfence_system();
              // (the value is a device timestamp in ~us, never 0: DFX_HOST_PIPE_TRACE prints it next to the host's clock)
              unsigned long long gt;
              asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
              *(volatile unsigned *)(p.pipe_host_flags + c) = (unsigned)(gt >> 10) | 1u;
            }
          }
        }
        if constexpr (EXTRA) {
          if (p.state_out != nullptr) {
            R *so = p.state_out + idx * (5 + D);
            so[0] = pid_inv; so[1] = pid_prev_inv; so[2] = at_dtmin ? R(1) : R(0); so[3] = made_jump ? R(1) : R(0);
            // first_step stays True only while no step has been kept (runge_kutta.py:1199-1200 sets it False on a kept step)
            bool first = FSAL && num_accepted == 0;
            if (first && p.state_in != nullptr && (p.state_in_flags & 2)) first = p.state_in[idx * (5 + D) + 4] != R(0);
            so[4] = first ? R(1) : R(0);
#pragma unroll
            for (int c = 0; c < D; ++c) so[5 + c] = FSAL ? f_fsal[c] : R(0);
          }
        }
        active = false;
      }
      if (p.totals != nullptr && waiting != 0u) {
        // ensemble totals for the multi-GPU statistics reduction: one warp reduction per finalise
        // pass into this warp's shared-memory slot; flushed with four global atomics when the warp leaves the kernel
        const bool fin = (waiting >> lane_id) & 1u;
        const int a_ = __reduce_add_sync(kFullMask, fin ? num_steps : 0), b_ = __reduce_add_sync(kFullMask, fin ? num_accepted : 0);
        // failed = not is_okay(result): an event is not a failure (_solution.py:52-62)
        const int c_ = __reduce_add_sync(kFullMask, (fin && result != DFX_RESULT_SUCCESSFUL && result != DFX_RESULT_EVENT_OCCURRED) ? 1 : 0);
        const int m_ = __reduce_max_sync(kFullMask, fin ? num_steps : 0);
        if (lane_id == 0) {
          long long w0, w1, w2, w3;
          asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "r"(tot_addr) : "memory");
          asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(w2), "=l"(w3) : "r"(tot_addr + 16u) : "memory");
          w0 += a_; w1 += b_; w2 += c_; w3 = w3 > m_ ? w3 : m_;
          asm volatile("st.shared.v2.u64 [%0], {%1, %2};" :: "r"(tot_addr), "l"(w0), "l"(w1) : "memory");
          asm volatile("st.shared.v2.u64 [%0], {%1, %2};" :: "r"(tot_addr + 16u), "l"(w2), "l"(w3) : "memory");
        }
      }
      if constexpr (RICH) {
        // Unfilled output slots read +inf (_integrate.py:1296-1300, 1320-1322).  The tails are written here, by the whole
        // warp per finished trajectory (coalesced streaming stores), instead of by a second kernel over the buffers: the
        // solve itself leaves most of the HBM write bandwidth idle, so the padding rides along for free (C3: the separate
        // pass cost 10.8 ms on top of a 22.8 ms solve) and every output byte is still written exactly once.
        const bool pad_saves = (p.save_ts != nullptr) || (p.save_steps > 0);
        const bool pad_dense = p.save_dense && !p.dense_lazy;
        if (pad_saves || pad_dense) {
          const int lane = lane_id;
          for (unsigned m = waiting; m; m &= m - 1) {
            const int src = __ffs(m) - 1;
            const long long r = __shfl_sync(kFullMask, idx_i, src);
            if (pad_saves) {
              const long long sc = __shfl_sync(kFullMask, save_index, src);
              pad_tail(p.ts_out + r * p.out_size, sc, (long long)p.out_size, lane, p.pad_vec != 0);
              pad_tail(p.ys_out + r * p.out_size * D, sc * D, (long long)p.out_size * D, lane, p.pad_vec != 0);
            }
            if (pad_dense) {
              const long long dc = __shfl_sync(kFullMask, dense_index, src), ms = p.max_steps;
              pad_tail(p.dense_ts + r * (ms + 1), dc + 1, ms + 1, lane, p.pad_vec != 0);
              pad_tail(p.dense_y0 + r * ms * D, dc * D, ms * D, lane, p.pad_vec != 0);
              pad_tail(p.dense_y1 + r * ms * D, dc * D, ms * D, lane, p.pad_vec != 0);
              if (DENSE_K && p.dense_k != nullptr) pad_tail(p.dense_k + r * ms * (S * D), dc * (S * D), ms * (S * D), lane, p.pad_vec != 0);
            }
          }
        }
      }
      // ---------------- refill: finished lanes claim the next trajectory ----------------
      // `exhausted` is warp-uniform (it is set from a warp vote), so the collective below is convergent.
      if (!exhausted) {
      const long long got = claim_work(!active, p.work_counter, lane_id);
      // the queue only grows: once any lane is handed an index past the end, it is drained for good
      exhausted = __any_sync(kFullMask, !active && got >= p.n_traj);
      if (!active) {
        if (got >= 0 && got < p.n_traj) {
          idx_i = (int)got;
          const long long idx = got;
          if constexpr (!RICH) {
            if (p.pipe_in_ready != nullptr) {  // wait until the copy engine has delivered this trajectory's chunk
              const unsigned need = (unsigned)(idx / p.pipe_chunk_len) + 1u;
              // pipe_in_ready[1] is the abort word (device memory, next to in_ready): the host sets it when it cannot
              // deliver the remaining chunks (e.g. a failed H2D enqueue) - the kernel must never outwait a host that gave up
              while (*(volatile const unsigned *)p.pipe_in_ready < need) {
                if (((volatile const unsigned *)p.pipe_in_ready)[1] != 0u) { pipe_aborted = true; break; }
                __nanosleep(200);
              }
              __threadfence();
            }
          }
          if (!pipe_aborted) {
          // _integrate.py:1076-1079, 1157-1165: time dtype and direction normalisation
          R a = p.t0_arr ? p.t0_arr[idx] : p.t0;
          R b = p.t1_arr ? p.t1_arr[idx] : p.t1;
          direction = (a < b) ? R(1) : R(-1);
          t0 = a * direction;
          t1 = b * direction;
#pragma unroll
          for (int c = 0; c < D; ++c) y[c] = p.y0[idx * D + c];
          if constexpr (PerTrajArgs<Field>::value) {
            if (p.traj_args != nullptr) {
#pragma unroll
              for (int i = 0; i < Field::kNumParams; ++i) fp_lane.p[i] = p.traj_args[idx * Field::kNumParams + i];
            }
          }
          // controller init: pid.py:316-392 (dt0=None -> 0.01) / constant.py:30-55
          R dt0 = p.has_dt0 ? p.dt0 * direction : R(0.01);
          if (p.controller == DFX_CTRL_PID) {
            if (p.has_dtmax) dt0 = jnp_min(dt0, p.dtmax);
            if (p.has_dtmin) dt0 = jnp_max(dt0, p.dtmin);
          } else {
            const R dt0_up = Num<R>::from_bits(Num<R>::bits(dt0) + (dt0 > R(0) ? 1 : (dt0 < R(0) ? -1 : 1)));  // nextafter(dt0, +inf)
            // constant.py:52-54: an infinite t1 (steady-state set-ups) is marked num_steps = -1 and steps by dt0
            cs_num_steps = r_isinf(t1) ? -1 : (int)ceil((double)((t1 - t0) / dt0_up));
          }
          tprev = t0;
          tnext = t0 + dt0;
          if constexpr (EXTRA) {  // ClipStepSizeController.init, clip.py:246-302
            made_jump = false;
            reject_index = p.n_reject;  // clip.py:292-299
            if (p.step_ts != nullptr) {
              step_index = clip_find_idx(t0, p.step_ts, p.n_step_ts, 0, direction);  // searchsorted(side="right")
              tnext = jnp_min(clip_get_t(p.step_ts, p.n_step_ts, step_index, direction), tnext);
            }
            if (p.jump_ts != nullptr) {
              jump_index = clip_find_idx(t0, p.jump_ts, p.n_jump_ts, 0, direction);
              tnext = jnp_min(prev_n<R>(clip_get_t(p.jump_ts, p.n_jump_ts, jump_index, direction), 1), tnext);
            }
          }
          tnext = jnp_min(tnext, t1);  // _integrate.py:1265
          t1_clip_floor = prev_n<R>(t1, 100);  // _integrate.py:320-322
          pid_inv = R(1); pid_prev_inv = R(1); at_dtmin = false;
          num_steps = 0; num_accepted = 0; result = DFX_RESULT_SUCCESSFUL;
          save_index = 0; saveat_ts_index = 0; dense_index = 0;
          if constexpr (SDE) bm.init(p.keys + 2 * idx, p.vbt);  // leaf key = split_by_tree(key, shape)[0] (tree.py:301)
          if constexpr (FSAL) {
            // runge_kutta.py:684-695: the first step evaluates stage 0 at (t0, y0); a rejected first
            // step re-evaluates the same point, so computing it once here is value-identical.
            Field::template eval<R>(fp, t0 * direction, y, f_fsal);
          }
          if constexpr (EXTRA) {  // resumed states (_integrate.py:1250-1271)
            if (p.state_in != nullptr) {
              const R *si = p.state_in + idx * (5 + D);
              if (p.state_in_flags & 1) { pid_inv = si[0]; pid_prev_inv = si[1]; at_dtmin = si[2] != R(0); }
              if (p.state_in_flags & 4) made_jump = si[3] != R(0);
              if constexpr (FSAL) {
                if ((p.state_in_flags & 2) && si[4] == R(0)) {  // (first_step == True means "evaluate", which was just done)
#pragma unroll
                  for (int c = 0; c < D; ++c) f_fsal[c] = si[5 + c];
                }
              }
            }
          }
          if constexpr (TAB && !SDE && EXTRA) {  // (the launcher routes hairer solves to the EXTRA instantiation)
            if (p.hairer && !p.has_dt0 && p.controller == DFX_CTRL_PID) {
              // _select_initial_step, pid.py:51-81 (Hairer, Norsett, Wanner II.4) - behind a flag: through diffeqsolve
              // the reference never reaches it.  func == terms.vf: WrapTerm passes t * direction.
              R f0[D], f1[D], ya[D], sc[D];
              if constexpr (FSAL) {
#pragma unroll
                for (int c = 0; c < D; ++c) f0[c] = f_fsal[c];
              } else {
                Field::template eval<R>(fp, t0 * direction, y, f0);
              }
              R s0 = R(0), s1 = R(0), s2 = R(0);
#pragma unroll
              for (int c = 0; c < D; ++c) {
                sc[c] = p.atol + r_abs(y[c]) * p.rtol;
                const R a = y[c] / sc[c], b = f0[c] / sc[c];
                s0 += a * a; s1 += b * b;
              }
              const R d0 = (D == 1) ? r_sqrt(s0) : r_sqrt(s0) / sqrt_d, d1 = (D == 1) ? r_sqrt(s1) : r_sqrt(s1) / sqrt_d;
              const bool cond = (d0 < R(1e-5)) || (d1 < R(1e-5));
              const R h0 = cond ? R(1e-6) : R(0.01) * (d0 / (cond ? R(1) : d1));
#pragma unroll
              for (int c = 0; c < D; ++c) ya[c] = y[c] + h0 * f0[c];
              Field::template eval<R>(fp, (t0 + h0) * direction, ya, f1);
#pragma unroll
              for (int c = 0; c < D; ++c) { const R a = (f1[c] - f0[c]) / sc[c]; s2 += a * a; }
              const R d2 = ((D == 1) ? r_sqrt(s2) : r_sqrt(s2) / sqrt_d) / h0;
              const R max_d = jnp_max(d1, d2);
              const R h1 = (max_d <= R(1e-15)) ? jnp_max(R(1e-6), h0 * R(1e-3)) : r_pow(R(0.01) / max_d, p.inv_error_order);
              R dth = jnp_min(R(100) * h0, h1);
              if (p.has_dtmax) dth = jnp_min(dth, p.dtmax);
              if (p.has_dtmin) dth = jnp_max(dth, p.dtmin);
              tnext = jnp_min(t0 + dth, t1);
            }
          }
          if constexpr (RICH) {
            if (p.save_dense) p.dense_ts[idx * (long long)(p.max_steps + 1)] = t0;  // 324-327; dense_ts stays in normalised time (DenseInterpolation applies `direction`)
            if (p.save_t0) {  // _integrate.py:329-341
              p.ts_out[idx * (long long)p.out_size] = t0 * direction;
#pragma unroll
              for (int c = 0; c < D; ++c) p.ys_out[(idx * (long long)p.out_size) * D + c] = y[c];
              save_index = 1;
            }
          }
          if constexpr (EXTRA) {
            for (int i = 0; i < p.n_events; ++i) event_value[i] = event_cond(i, tprev, y, direction);  // _integrate.py:1432-1476
          }
          active = true;
          runnable = (tprev < t1) && (0 < p.max_steps) && (SPEC || result == DFX_RESULT_SUCCESSFUL);
          }
        }
      }
      if constexpr (!RICH) {
        if (p.pipe_in_ready != nullptr && __any_sync(kFullMask, pipe_aborted)) exhausted = true;  // stop claiming
      }
      }
    }
    if (__all_sync(kFullMask, !active)) break;

    // ---------------- one attempted step (all active lanes, same instruction stream) ----------------
    [[maybe_unused]] long long dense_row = -1;  // >= 0: this lane staged a dense record in shared memory this iteration
    if (active) {
      const bool run = runnable;  // 355-363, 685-687: tprev < t1, num_steps < max_steps, result == successful
      if (run) {
        [[maybe_unused]] const long long idx = idx_i;
        const R st0 = tprev, st1 = tnext;
        const R dt = st1 - st0;
        // One step of the (unwrapped) solver on [st0, st1] from y: AbstractRungeKutta.step / Euler.step / the ShARK branch
        // of AbstractSRK.step.  f_in = carried FSAL derivative, reeval = made_jump; k, y1, yerr, f_last are the outputs.
        auto single_step = [&](const R st0, const R st1, const R (&y)[D], [[maybe_unused]] const R (&f_in)[D],
                               [[maybe_unused]] const bool reeval, R (&y1)[D], R (&yerr)[D], R (&k)[S][D], R (&f_last)[D]) {
          const R dt = st1 - st0;
          // one Brownian query per step and component (runge_kutta.py:644, srk.py:390); Wc(c) / Hc(c): the increment driving component c
          R Wv[NW], Hv[NW];
#pragma unroll
          for (int w = 0; w < NW; ++w) { Wv[w] = R(0); Hv[w] = R(0); }
          if constexpr (SDE) bm.increment(st0, st1, p.vbt, Wv, Hv);
          auto Wc = [&](int c) { return Wv[NW == 1 ? 0 : (c < NW ? c : 0)]; };
          auto Hc = [&](int c) { return Hv[NW == 1 ? 0 : (c < NW ? c : 0)]; };
          // ControlTerm.prod (_term.py:417-427): g(t) . X for component c, X = W or H.  Scalar / diagonal diffusion: g(t) X_c;
          // matrix diffusion: tensordot over the Brownian axis
          [[maybe_unused]] auto gprod = [&](R t, const R (&yy)[D], const R (&X)[NW], int c) -> R {
            if constexpr (!SDE) return R(0);
            else if constexpr (StateNoise<Field>::value) return Field::template noise_prod<R>(fp, t, yy, X, c);  // g(t, y) . X
            else if constexpr (MatrixNoise<Field>::value) return Field::template noise_prod<R>(fp, t, X, c);
            else return Field::template diffusion<R>(fp, t) * X[NW == 1 ? 0 : (c < NW ? c : 0)];
          };

          if constexpr (TAB) {
            // ---- explicit RK step, runge_kutta.py:643-1203 ----
            const R control = direction * dt;  // WrapTerm.contr (_term.py:742-745)
            R yi[D], fi[D];
            if constexpr (FSAL) {
              R f0[D];
  #pragma unroll
              for (int c = 0; c < D; ++c) f0[c] = f_in[c];
              if constexpr (EXTRA) {
                // eval_first_stage = first_step | made_jump (runge_kutta.py:687): after stepping around a jump the carried
                // derivative belongs to the other side of the discontinuity and is re-evaluated at (st0, y)
                if (reeval) Field::template eval<R>(fp, st0 * direction, y, f0);
              }
  #pragma unroll
              for (int c = 0; c < D; ++c) k[0][c] = control * f0[c];  // prod(f0), 695 & 771
            } else {
              Field::template eval<R>(fp, st0 * direction, y, fi);
  #pragma unroll
              for (int c = 0; c < D; ++c) {
                R kk = control * fi[c];
                if constexpr (SDE) kk = kk + gprod(st0, y, Wv, c);  // MultiTerm.vf_prod (_term.py:711-722)
                k[0][c] = kk;
              }
            }
  #pragma unroll
            for (int c = 0; c < D; ++c) yi[c] = y[c];
  #pragma unroll
            for (int i = 1; i < S; ++i) {  // rk_stage, 847-1069
  #pragma unroll
              for (int c = 0; c < D; ++c) {
                if constexpr (kChainY0) {
                  // y_i = y0 + sum_j a_ij k_j as ONE chain of FMAs seeded with y0: saves the separate add per stage and
                  // component (18 of ~200 FP64 instructions of a Lorenz/Dopri5 step) at the price of rounding each partial
                  // sum at ulp(y0) instead of once
                  R acc = y[c];
  #pragma unroll
                  for (int j = 0; j < i; ++j)
                    if (Solver::hA(i * (i - 1) / 2 + j) != 0.0) acc += Solver::template a<R>(i, j) * k[j][c];
                  yi[c] = acc;
                } else {
                  R incr = R(0);
  #pragma unroll
                  for (int j = 0; j < i; ++j)
                    if (Solver::hA(i * (i - 1) / 2 + j) != 0.0) incr += Solver::template a<R>(i, j) * k[j][c];  // vector_tree_dot (base.py:37-41); structural zeros skipped
                  yi[c] = y[c] + incr;  // 871
                }
              }
              const R ti = (Solver::hC(i) == 1.0) ? st1 : st0 + Solver::template c<R>(i) * dt;  // 1023
              Field::template eval<R>(fp, ti * direction, yi, fi);
  #pragma unroll
              for (int c = 0; c < D; ++c) {
                if constexpr (kLastStageF && !RICH) {
                  if (i == S - 1) continue;  // the last stage value is only read by the error estimate, through f_last below
                }
                R kk = control * fi[c];
                if constexpr (SDE) kk = kk + gprod(ti, yi, Wv, c);
                k[i][c] = kk;
              }
            }
  #pragma unroll
            for (int c = 0; c < D; ++c) f_last[c] = fi[c];
            if constexpr (Solver::kSsal) {  // 1161
  #pragma unroll
              for (int c = 0; c < D; ++c) y1[c] = yi[c];
            } else {  // 1177-1185
  #pragma unroll
              for (int c = 0; c < D; ++c) {
                R incr = R(0);
  #pragma unroll
                for (int j = 0; j < S; ++j)
                  if (Solver::hBsol(j) != 0.0) incr += Solver::template b_sol<R>(j) * k[j][c];
                y1[c] = y[c] + incr;
              }
            }
  #pragma unroll
            for (int c = 0; c < D; ++c) {  // 1186-1193
              R e = R(0);
              if constexpr (kLastStageF) {
                // FSAL+SSAL pairs: k_{s-1} = dt f(y1) enters only here, so take it as (b_err[s-1] dt) f(y1) and keep f(y1)
                // (the next step's FSAL derivative) as the one live copy instead of k_{s-1} AND f(y1)
  #pragma unroll
                for (int j = 0; j < S - 1; ++j)
                  if (Solver::hBerr(j) != 0.0) e += Solver::template b_err<R>(j) * k[j][c];
                if (Solver::hBerr(S - 1) != 0.0) e += (Solver::template b_err<R>(S - 1) * control) * f_last[c];
              } else {
  #pragma unroll
                for (int j = 0; j < S; ++j)
                  if (Solver::hBerr(j) != 0.0) e += Solver::template b_err<R>(j) * k[j][c];
              }
              yerr[c] = e;
            }
          } else if constexpr (InnerId<Solver>::value == DFX_EULER) {
            // ---- euler.py:46-59 ----
            R f0[D];
            Field::template eval<R>(fp, st0 * direction, y, f0);
  #pragma unroll
            for (int c = 0; c < D; ++c) {
              R kk = (direction * dt) * f0[c];
              if constexpr (SDE) kk = kk + gprod(st0, y, Wv, c);
              k[0][c] = kk;
              y1[c] = y[c] + kk;
              yerr[c] = R(0);
              f_last[c] = R(0);
            }
          } else {
            // ---- ShARK: srk.py:335-671 additive-noise branch with shark.py:10-30 ----
            if constexpr (SDE) {
              const R h = dt;
              // w_kg = g(t0) . W, h_kg = g(t0) . H (441-447); g_delta = (g(t1) - g(t0)) / 2 applied to W - 2 H (612-618)
              auto w_kg = [&](int c) { return gprod(st0, y, Wv, c); };
              auto h_kg = [&](int c) { return gprod(st0, y, Hv, c); };
              auto time_var = [&](int c) -> R {
                if constexpr (MatrixNoise<Field>::value) {  // constant matrix: g_delta = 0.5 (G - G) = 0 exactly, so prod(g_delta, .) = 0
                  return R(0);
                } else {
                  const R g0 = Field::template diffusion<R>(fp, st0), g1 = Field::template diffusion<R>(fp, st1);
                  return (R(0.5) * (g1 - g0)) * (Wc(c) - R(2.0) * Hc(c));
                }
              };
              R z[D], fz[D];
  #pragma unroll
              for (int c = 0; c < D; ++c) z[c] = y[c] + R(0) + (R(kSharkAW0) * w_kg(c) + R(kSharkAH0) * h_kg(c));  // stage 0: 545
              Field::template eval<R>(fp, st0, z, fz);  // 548: t0 + 0*h
  #pragma unroll
              for (int c = 0; c < D; ++c) k[0][c] = h * fz[c];
  #pragma unroll
              for (int c = 0; c < D; ++c) z[c] = y[c] + R(kSharkA10) * k[0][c] + (R(kSharkAW1) * w_kg(c) + R(kSharkAH1) * h_kg(c));
              Field::template eval<R>(fp, st0 + R(kSharkC1) * h, z, fz);
  #pragma unroll
              for (int c = 0; c < D; ++c) k[1][c] = h * fz[c];
  #pragma unroll
              for (int c = 0; c < D; ++c) {
                R diffusion_result = R(kSharkBW) * w_kg(c) + R(kSharkBH) * h_kg(c);   // 603-607
                diffusion_result = diffusion_result + time_var(c);  // 612-618
                yerr[c] = R(kSharkE0) * k[0][c] + R(kSharkE1) * k[1][c];        // 638-639, 663
                const R drift_result = R(kSharkB0) * k[0][c] + R(kSharkB1) * k[1][c];  // 667
                y1[c] = y[c] + drift_result + diffusion_result;                  // 669
                f_last[c] = R(0);
              }
            }
          }
        };
        static_assert(!(EXTRA && !RICH), "EXTRA implies RICH");
        R k[S][D];
        R y1[D], yerr[D], f_last[D];
        [[maybe_unused]] R y1_alt[D];
        if constexpr (!IsHalf<Solver>::value) {
          single_step(st0, st1, y, f_fsal, made_jump, y1, yerr, k, f_last);
        } else {  // HalfSolver.step, base.py:312-341
          R yhalf[D], f_half[D], f_alt[D], e_[D], k_half[S][D];
          const R thalf = st0 + R(0.5) * (st1 - st0);
          single_step(st0, thalf, y, f_fsal, made_jump, yhalf, e_, k_half, f_half);
          single_step(thalf, st1, yhalf, f_half, false, y1, e_, k_half, f_last);
          single_step(st0, st1, y, f_fsal, made_jump, y1_alt, e_, k, f_alt);  // dense_info comes from the full step
#pragma unroll
          for (int c = 0; c < D; ++c) yerr[c] = r_abs(y1[c] - y1_alt[c]);
        }
        // the end point the local interpolant / dense output sees: dense_info["y1"] (y1_alt under HalfSolver)
        [[maybe_unused]] auto &y1_dense = *(IsHalf<Solver>::value ? &y1_alt : &y1);

        // ---- step-size controller ----
        bool keep;
        R next_t0, next_t1;
        if (SPEC || p.controller == DFX_CTRL_PID) {
          // pid.py:394-567.  y_error NaN -> inf first (_integrate.py:386).
          R dtn, inv = R(1), factor;
          bool slow = true;
          if constexpr (FAST_PID) {
            // ---- fast path: pure I-controller (pcoeff = dcoeff = 0, icoeff = 1) at the solver's own order ----
            // scaled_error = sqrt(ss / D), so   keep <=> ss/D < 1   and   factor = safety * (ss/D)^(-1/(2 order)).
            // No sqrt, no IEEE division, no pow(), no FP64 compares: reciprocals and the 2*order-th root are
            // Newton-refined SFU seeds, max/min/clip run on the integer ALU, and the range / accept tests are made
            // on (float)q.  NaN or inf anywhere (y1, y_error) makes q non-finite and fails the range test, and q
            // within 1e-6 of the accept boundary (where the reference's own rounding of sqrt and '/' decides) is
            // excluded too: all of those take the faithful path below, so accept/reject decisions are the reference's.
            if (SPEC || p.fast_pid) {
              R ss = R(0);
#pragma unroll
              for (int c = 0; c < D; ++c) {  // _scale, 483-490 (a NaN y1 propagates through abs_max_bits)
#if DFX_OPT_ABSMAX_FP64
                const R ya_ = r_abs(y[c]), yb_ = r_abs(y1[c]);
                const R yy = (ya_ > yb_) ? ya_ : yb_;  // (a NaN y1 fails the compare and is selected: it propagates)
#else
                const R yy = abs_max_bits(y[c], y1[c]);
#endif
                const R sc = yerr[c] * fast_rcp(p.atol + yy * p.rtol);
                ss += sc * sc;
              }
              const R q = ss * R(1.0 / D);
              const float qf = (float)q;
              if (qf > 1e-30f && qf < 1e30f && fabsf(qf - 1.0f) > 1e-6f) {
                slow = false;
                keep = qf < 1.0f;                                  // 493 (same decision as q < 1 outside the 1e-6 band)
                if (!SPEC && p.has_dtmin) keep = keep || at_dtmin;          // 495-496
                factor = p.safety * (R)inv_root<2 * Solver::kOrder>((double)q, qf);  // 515, 522
                const R fmin = keep ? R(1) : p.factormin;          // 518
                const R fmax = keep ? p.factormax : p.safety;      // 520
                factor = pos_min_bits(pos_max_bits(factor, fmin), fmax);  // 521-525 (all operands positive)
                dtn = x_mul(dt, factor);                           // 531 (own rounding: never fused into next_t0 + dtn)
              }
            }
          }
          if (slow) {
            bool nan_any = false;
#pragma unroll
            for (int c = 0; c < D; ++c) nan_any |= r_isnan(y1[c]);
            R ss = R(0), sc0 = R(0);
#pragma unroll
            for (int c = 0; c < D; ++c) {  // _scale, 483-490
              const R e = r_isnan(yerr[c]) ? Num<R>::inf() : yerr[c];
              const R yc = nan_any ? y[c] : y1[c];
              const R yy = r_max(r_abs(y[c]), r_abs(yc));
              const R sc = e / (p.atol + yy * p.rtol);
              ss += sc * sc;
              sc0 = sc;
            }
            const R scaled_error = (D == 1) ? r_abs(sc0) : r_sqrt(ss) / sqrt_d;  // optx.rms_norm
            keep = scaled_error < R(1);                     // 493
            if (!SPEC && p.has_dtmin) keep = keep || at_dtmin;       // 495-496
            inv = R(1) / scaled_error;                      // 498
            factor = p.safety;
            if (p.use_c1) factor = factor * r_pow(inv, p.coeff1);          // 515
            if (!SPEC && p.use_c2) factor = factor * r_pow(pid_inv, p.coeff2);      // 516 (SPEC: pure I-controller, no history)
            if (!SPEC && p.use_c3) factor = factor * r_pow(pid_prev_inv, p.coeff3); // 517
            const R fmin = keep ? R(1) : p.factormin;       // 518
            const R fmax = keep ? p.factormax : p.safety;   // 520
            factor = jnp_min(jnp_max(factor, fmin), fmax);  // 521-525
            dtn = x_mul(dt, factor);                        // 531
            if (inv == R(0) || r_isinf(inv)) inv = R(1);    // 537-538
          }
          if (!SPEC && (p.has_dtmax | p.has_dtmin)) {  // one uniform branch around both limits: neither is set by default
            if (p.has_dtmax) dtn = jnp_min(dtn, p.dtmax);   // 545-546
            if (p.has_dtmin) {                              // 547-555
              if (!p.force_dtmin && dtn < p.dtmin && result == DFX_RESULT_SUCCESSFUL) result = DFX_RESULT_DT_MIN_REACHED;
              if (at_dtmin && factor == R(1)) dtn = p.dtmin;
              at_dtmin = dtn <= p.dtmin;
              dtn = jnp_max(dtn, p.dtmin);
            }
          }
          next_t0 = keep ? st1 : st0;                     // 557-558
          next_t1 = x_add(next_t0, dtn);  // the reference's two roundings, identical in every instantiation of the kernel
          if constexpr (!SPEC) { if (keep) { pid_prev_inv = pid_inv; pid_inv = inv; } }  // 560-564 (unused by the pure I-controller)
        } else {
          // constant.py:57-104
          keep = true;
          const int cs_steps_completed = num_steps + 2;
          R t1n = t0 + (t1 - t0) * ((R)cs_steps_completed / (R)cs_num_steps);
          if (cs_steps_completed == cs_num_steps) t1n = t1;
          if (cs_num_steps < 0) t1n = st1 + p.dt0 * direction;  // constant.py:93 (t1_sim_or_dt0 = dt0)
          next_t0 = st1;
          next_t1 = t1n;
        }

        if constexpr (EXTRA) {  // ClipStepSizeController.adapt_step_size, clip.py:350-377
          bool ctrl_made_jump = false;
          if (p.step_ts != nullptr) {
            bool dummy;
            const R nt0 = clip_bump(next_t0, p.step_ts, p.n_step_ts, direction, dummy);
            step_index = clip_find_idx(nt0, p.step_ts, p.n_step_ts, step_index, direction);
            next_t1 = jnp_min(clip_get_t(p.step_ts, p.n_step_ts, step_index, direction), next_t1);
          }
          if (p.jump_ts != nullptr) {
            next_t0 = clip_bump(next_t0, p.jump_ts, p.n_jump_ts, direction, ctrl_made_jump);
            next_t1 = jnp_max(next_after_up(next_t0), next_t1);
            jump_index = clip_find_idx(next_t0, p.jump_ts, p.n_jump_ts, jump_index, direction);
            next_t1 = jnp_min(prev_n<R>(clip_get_t(p.jump_ts, p.n_jump_ts, jump_index, direction), 1), next_t1);
          }
          if (p.reject_ts != nullptr) {  // clip.py:398-424 (t1 there is the attempted step's end st1)
            R *rts = p.reject_ts + idx * (long long)p.n_reject;
            const R rejected_t = (reject_index == p.n_reject) ? Num<R>::inf() : rts[reject_index];
            if (st1 > rejected_t && result == DFX_RESULT_SUCCESSFUL) result = DFX_RESULT_INTERNAL_ERROR;
            reject_index += (st1 == rejected_t) ? 1 : 0;
            reject_index -= keep ? 0 : 1;
            if (reject_index < 0 && result == DFX_RESULT_SUCCESSFUL) result = DFX_RESULT_MAX_STEPS_REJECTED;
            if (reject_index >= 0 && reject_index < p.n_reject && !keep) rts[reject_index] = st1;
            const R clip_to = (reject_index >= p.n_reject) ? Num<R>::inf() : rts[reject_index < 0 ? p.n_reject - 1 : reject_index];
            next_t1 = jnp_min(clip_to, next_t1);
          }
          if (keep) made_jump = ctrl_made_jump;  // _integrate.py:425
        }

        // ---- book-keeping, _integrate.py:412-437 ----
        // 412: tprev = min(tprev, t1) is the identity without jump_ts: next_t0 is st0 or st1, and tnext never exceeds t1
        // (it starts as min(t0 + dt0, t1) and every update below clips it to t1); a bumped next_t0 is clipped explicitly.
        const R tprev_new = (EXTRA && p.jump_ts != nullptr) ? jnp_min(next_t0, t1) : next_t0;
        R tnext_new = next_t1;
        if (next_t1 > t1_clip_floor) tnext_new = keep ? t1 : tprev_new + R(0.5) * (t1 - tprev_new);  // 278-284
        num_steps += 1;
        num_accepted += keep ? 1 : 0;

        // ---- Event: _integrate.py:548-633 (detection at the new state of every step) and 691-821 (root find) ----
        // The reference saves the step, then finds the event time, then deletes every saved time after it ("unsave");
        // here the event time is found first and the saves below simply stop at it - the same buffers come out.
        [[maybe_unused]] bool ev_hit = false;
        [[maybe_unused]] R t_event = tprev_new;
        [[maybe_unused]] R y_event[D];
        if constexpr (EXTRA) {
          if (p.n_events != 0) {
            R ynew[D];
#pragma unroll
            for (int c = 0; c < D; ++c) ynew[c] = keep ? y1[c] : y[c];
            // every condition is re-evaluated; the first one (in PyTree order) that triggers decides (619-626)
            bool m = false;
            int which = 0;
            for (int i = 0; i < p.n_events; ++i) {
              const R nv = event_cond(i, tprev_new, ynew, direction);
              const int so = (event_value[i] > R(0)) - (event_value[i] < R(0)), sn = (nv > R(0)) - (nv < R(0));
              bool mi;
              if (p.event_kind[i] == DFX_EVENT_STEADY_STATE) mi = nv != R(0);
              else if (p.event_dir[i] == 0) mi = so != sn;
              else if (p.event_dir[i] == 1) mi = (so <= 0) && (sn > 0);
              else mi = (so > 0) && (sn <= 0);
              event_value[i] = nv;
              if (mi && !m) { m = true; which = i; }
            }
            if (m) {
              ev_hit = true;
              result = DFX_RESULT_EVENT_OCCURRED;
              if (p.event_root) {
                R tf = st1;
                bool ok = true;
                if (p.event_kind[which] == DFX_EVENT_AFFINE || p.event_kind[which] == DFX_EVENT_USER) {
                  // [EXT] optimistix.Newton(rtol, atol), options lower / upper = the step, y0 = its end, max_steps 256:
                  // clipped Newton steps; Cauchy termination on the iterate and on the function value
                  const R fd_h = (st1 - st0) * (sizeof(R) == 8 ? R(1e-6) : R(1e-3));
                  auto along = [&](R t, R &g, R &dg) {
                    R yq[D], dq[D];
                    interp_eval<INTERP, R, S, D>(st0, st1, y, y1_dense, k, t, yq);
                    if (p.event_kind[which] == DFX_EVENT_USER) {
                      // the user's condition has no derivative to offer (the reference differentiates cond_fn by JVP): central
                      // difference of t -> cond(t, interpolant(t)) over 1e-6 of the step (the root itself only depends on cond)
                      g = event_cond(which, t, yq, direction);
                      R ya[D], yb[D];
                      interp_eval<INTERP, R, S, D>(st0, st1, y, y1_dense, k, t + fd_h, ya);
                      interp_eval<INTERP, R, S, D>(st0, st1, y, y1_dense, k, t - fd_h, yb);
                      dg = (event_cond(which, t + fd_h, ya, direction) - event_cond(which, t - fd_h, yb, direction)) / (R(2) * fd_h);
                      return;
                    }
                    interp_deriv<INTERP, R, S, D>(st0, st1, y, y1_dense, k, t, dq);
                    R v = R(0), dv = R(0);
#pragma unroll
                    for (int c = 0; c < D; ++c) { v += p.ev_w[which][c < 4 ? c : 3] * yq[c]; dv += p.ev_w[which][c < 4 ? c : 3] * dq[c]; }
                    g = v + p.ev_wt[which] * t + p.ev_b[which];
                    dg = dv + p.ev_wt[which];
                  };
                  R g, dg;
                  along(tf, g, dg);
                  ok = false;
                  for (int it = 0; it < 256; ++it) {
                    R tn = jnp_min(jnp_max(tf - g / dg, st0), st1);
                    R gn, dgn;
                    along(tn, gn, dgn);
                    const bool conv = (r_abs(tn - tf) < p.ev_atol + p.ev_rtol * r_abs(tn)) && (r_abs(gn - g) < p.ev_atol + p.ev_rtol * r_abs(gn));
                    tf = tn; g = gn; dg = dgn;
                    if (conv) { ok = true; break; }
                  }
                }
                t_event = tf;
                interp_eval<INTERP, R, S, D>(st0, st1, y, y1_dense, k, tf, y_event);
                if (!ok) result = DFX_RESULT_EVENT_ROOT_FIND_FAILED;
              }
            }
          }
        }
        // with a root finder, saves of this step at times after the event time do not survive (777-806)
        [[maybe_unused]] const bool ev_cut = EXTRA && ev_hit && p.event_root;

        if constexpr (RICH) {
          // ---- SaveAt(ts): interpolant on the attempted interval, kept steps only (456-487) ----
          if (p.save_ts != nullptr && keep) {
            while (saveat_ts_index < p.n_save_ts) {
              const R tq = p.save_ts[saveat_ts_index] * direction;
              if (!(tq <= st1)) break;
              if (ev_cut && tq > t_event) break;
              R yq[D];
              interp_eval<INTERP, R, S, D>(st0, st1, y, y1_dense, k, tq, yq);
              const long long o = idx * (long long)p.out_size + save_index;
              p.ts_out[o] = tq * direction;  // final ts *= direction (1479-1482)
#pragma unroll
              for (int c = 0; c < D; ++c) p.ys_out[o * D + c] = yq[c];
              saveat_ts_index += 1;
              save_index += 1;
            }
          }
          // ---- SaveAt(steps=n) (493-524) ----
          if (p.save_steps != 0 && keep && (num_accepted % p.save_steps) == 0 && !(ev_cut && tprev_new > t_event)) {
            const long long o = idx * (long long)p.out_size + save_index;
            p.ts_out[o] = tprev_new * direction;
#pragma unroll
            for (int c = 0; c < D; ++c) p.ys_out[o * D + c] = y1[c];
            save_index += 1;
          }
          // ---- SaveAt(dense=True) (529-540) ----
          if (p.save_dense && keep) {
            const long long row = idx * (long long)p.max_steps + dense_index;
            if (p.dense_cs) st_cs(&p.dense_ts[idx * (long long)(p.max_steps + 1) + dense_index + 1], tprev_new);
            else p.dense_ts[idx * (long long)(p.max_steps + 1) + dense_index + 1] = tprev_new;
            if (p.dense_coop) {
              // stage this lane's record {k[S][D], y0[D], y1[D]} in shared memory; the warp flushes it below
              R *rec = dense_smem + ((threadIdx.x >> 5) * 32 + (threadIdx.x & 31)) * kDenseStride;
#pragma unroll
              for (int j = 0; j < S; ++j)
#pragma unroll
                for (int c = 0; c < D; ++c) rec[DENSE_K ? j * D + c : 0] = k[j][c];
#pragma unroll
              for (int c = 0; c < D; ++c) { rec[kDenseK + c] = y[c]; rec[kDenseK + D + c] = y1_dense[c]; }
              dense_row = row;
            } else {
              store_row<D>(&p.dense_y0[row * D], y, p.dense_vec_ok != 0, p.dense_cs != 0);
              store_row<D>(&p.dense_y1[row * D], y1_dense, p.dense_vec_ok != 0, p.dense_cs != 0);
              if constexpr (DENSE_K) {
                if (p.dense_k != nullptr) {
                  R flat[S * D];
#pragma unroll
                  for (int j = 0; j < S; ++j)
#pragma unroll
                    for (int c = 0; c < D; ++c) flat[j * D + c] = k[j][c];
                  store_row<S * D>(&p.dense_k[row * (S * D)], flat, p.dense_vec_ok != 0, p.dense_cs != 0);
                }
              }
            }
            dense_index += 1;
          }
        }

        // keep-or-restore (422-424)
#pragma unroll
        for (int c = 0; c < D; ++c) {
          y[c] = keep ? y1[c] : y[c];
          if constexpr (FSAL) f_fsal[c] = keep ? f_last[c] : f_fsal[c];
        }
        tprev = tprev_new;
        tnext = tnext_new;
        if constexpr (EXTRA) {
          if (ev_cut) {  // tfinal, yfinal = the event time and the interpolant there (745-756)
            tprev = t_event;
#pragma unroll
            for (int c = 0; c < D; ++c) y[c] = y_event[c];
          }
        }
      }
    }

    // ---------------- SaveAt(dense): warp-cooperative flush of the staged records ----------------
    // A lane's record is (S + 2) D contiguous values per output array row; written by its own lane it would be
    // (S + 2) D / 4 separate 32-byte sector writes scattered over 32 trajectories' rows per instruction.  Here the whole
    // warp writes one record at a time: consecutive lanes store consecutive elements, i.e. full contiguous lines.
    if constexpr (RICH) {
      if (p.save_dense && p.dense_coop) {
        const unsigned staged = __ballot_sync(kFullMask, dense_row >= 0);
        if (staged) {
          __syncwarp();
          const int lane = threadIdx.x & 31;
          const R *wrec = dense_smem + (threadIdx.x >> 5) * 32 * kDenseStride;
          if (kDenseVec && p.dense_vec_ok && p.flush_vec) {
            // kDenseGroup lanes per record, 32 / kDenseGroup records per pass: lane (g, q) moves vector q of the g-th
            // staged record of this pass
            const int q = lane & (kDenseGroup - 1), g = lane / kDenseGroup;
            unsigned m = staged;
            while (m) {
              int src = -1;
#pragma unroll
              for (int gg = 0; gg < 32 / kDenseGroup; ++gg) {  // hand the next 32 / kDenseGroup staged lanes to the groups
                const int s_ = m ? __ffs(m) - 1 : -1;
                if (m) m &= m - 1;
                if (gg == g) src = s_;
              }
              const long long row = __shfl_sync(kFullMask, dense_row, src < 0 ? 0 : src);
              if (src >= 0 && q < kDenseChunks) {
                const R *v = wrec + src * kDenseStride + q * kVW;
                R *dst;
                if (q < kDenseK / kVW) dst = p.dense_k + row * kDenseK + q * kVW;
                else if (q < (kDenseK + D) / kVW) dst = p.dense_y0 + row * D + (q * kVW - kDenseK);
                else dst = p.dense_y1 + row * D + (q * kVW - kDenseK - D);
                st32B(dst, v, p.dense_cs != 0);
              }
            }
          } else
          for (unsigned m = staged; m; m &= m - 1) {
            const int src = __ffs(m) - 1;
            const long long row = __shfl_sync(kFullMask, dense_row, src);
#pragma unroll
            for (int e0 = 0; e0 < kDenseRec; e0 += 32) {
              const int e = e0 + lane;
              if (e < kDenseRec) {
                const R v = wrec[src * kDenseStride + e];
                R *dst;
                if (e < kDenseK) dst = p.dense_k + row * kDenseK + e;
                else if (e < kDenseK + D) dst = p.dense_y0 + row * D + (e - kDenseK);
                else dst = p.dense_y1 + row * D + (e - kDenseK - D);
                if (p.dense_cs) st_cs(dst, v); else *dst = v;
              }
            }
          }
          __syncwarp();
        }
      }
    }

  }
  if (p.totals != nullptr && lane_id == 0) {
    long long w0, w1, w2, w3;
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "r"(tot_addr) : "memory");
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(w2), "=l"(w3) : "r"(tot_addr + 16u) : "memory");
    if (w0) atomicAdd((unsigned long long *)p.totals + 0, (unsigned long long)w0);
    if (w1) atomicAdd((unsigned long long *)p.totals + 1, (unsigned long long)w1);
    if (w2) atomicAdd((unsigned long long *)p.totals + 2, (unsigned long long)w2);
    if (w3) atomicMax(p.totals + 3, w3);
  }
}

}  // namespace dfx
