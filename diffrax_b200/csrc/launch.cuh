// launch.cuh - host-side launcher template for ensemble_kernel and the launcher registry.
#pragma once
#include <cstdlib>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "ensemble_kernel.cuh"

namespace dfx {

// thread-local error string + launch counter (api.cu)
void set_error(const char *fmt, ...);
void count_launch(int n = 1);
int register_builtin(int field_id, int dim, int solver_id, int dtype, int levy, dfx_launcher_fn fn);

// Host-pipelined mode: dfx_ensemble_solve_host parks the control words of its chunk pipeline here (thread-local,
// api.cu) around its one dfx_ensemble_solve call; the launcher that picks them up says so through `consumed`.
struct HostPipe {
  const unsigned *in_ready;
  unsigned *done, *host_flags;
  int chunk_len;
  bool consumed;
};
HostPipe *&host_pipe();

#define DFX_CUDA_OK(expr)                                                                  \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      dfx::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DFX_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

inline bool is_nan(double x) { return x != x; }

// Fill the dtype-typed SolveParams from the descriptor.  Everything that the reference computes in
// Python floats (error order, PID exponents) is computed here in double and rounded once.
template <class R, class Solver>
int fill_params(const dfx_solve_desc *d, SolveParams<R> &p, bool sde) {
  std::memset(&p, 0, sizeof(p));
  p.n_traj = d->n_traj;
  p.y0 = (const R *)d->y0;
  p.t0_arr = (const R *)d->t0_per_traj;
  p.t1_arr = (const R *)d->t1_per_traj;
  p.t0 = (R)d->t0;
  p.t1 = (R)d->t1;
  p.has_dt0 = !is_nan(d->dt0);
  p.dt0 = p.has_dt0 ? (R)d->dt0 : (R)0;
  p.controller = d->controller;
  p.rtol = (R)d->rtol; p.atol = (R)d->atol; p.safety = (R)d->safety;
  p.factormin = (R)d->factormin; p.factormax = (R)d->factormax;
  p.has_dtmin = !is_nan(d->dtmin); p.has_dtmax = !is_nan(d->dtmax);
  p.dtmin = p.has_dtmin ? (R)d->dtmin : (R)0;
  p.dtmax = p.has_dtmax ? (R)d->dtmax : (R)0;
  p.force_dtmin = d->force_dtmin;
  // base.py:97-120: ODE -> order ; SDE -> strong_order + 0.5 (Euler/Heun 0.5, ShARK 1.5)
  double error_order;
  if (!is_nan(d->error_order)) error_order = d->error_order;
  else if (sde) error_order = (InnerId<Solver>::value == DFX_SHARK ? 1.5 : 0.5) + 0.5;  // HalfSolver: the same (base.py:291-295)
  else error_order = (double)Solver::kOrder;
  const double c1 = (d->icoeff + d->pcoeff + d->dcoeff) / error_order;  // pid.py:512-514
  const double c2 = -(d->pcoeff + 2 * d->dcoeff) / error_order;
  const double c3 = d->dcoeff / error_order;
  p.coeff1 = (R)c1; p.coeff2 = (R)c2; p.coeff3 = (R)c3;
  p.use_c1 = c1 != 0; p.use_c2 = c2 != 0; p.use_c3 = c3 != 0;
  p.step_ts = (const R *)d->step_ts; p.n_step_ts = d->step_ts ? d->n_step_ts : 0;
  p.jump_ts = (const R *)d->jump_ts; p.n_jump_ts = d->jump_ts ? d->n_jump_ts : 0;
  p.hairer = d->hairer_initial_step && !sde;
  p.inv_error_order = (R)(1.0 / error_order);
  p.fast_pid = !sde && d->pcoeff == 0 && d->dcoeff == 0 && d->icoeff == 1 && error_order == (double)Solver::kOrder;
  p.save_t0 = d->save_t0; p.save_t1 = d->save_t1; p.save_steps = d->save_steps; p.save_dense = d->save_dense;
  p.save_ts = (const R *)d->save_ts;
  p.n_save_ts = d->save_ts ? d->n_save_ts : 0;
  p.max_steps = d->max_steps;
  p.out_size = dfx_out_size(d);
  p.ts_out = (R *)d->ts_out; p.ys_out = (R *)d->ys_out;
  p.stats = d->stats; p.result = d->result; p.save_count = d->save_count;
  p.dense_ts = (R *)d->dense_ts; p.dense_y0 = (R *)d->dense_y0; p.dense_y1 = (R *)d->dense_y1; p.dense_k = (R *)d->dense_k;
  p.dense_count = d->dense_count;
  p.dense_lazy = d->dense_lazy_padding;
  p.dense_vec_ok = (((uintptr_t)d->dense_y0 | (uintptr_t)d->dense_y1 | (uintptr_t)d->dense_k) & 31u) == 0;
  p.y_final = (R *)d->y_final; p.t_final = (R *)d->t_final;
  p.totals = (long long *)d->totals;
  p.n_peers = d->n_peers; p.peer_row0 = d->peer_row_offset;
  for (int q = 0; q < DFX_MAX_PEERS; ++q) { p.peer_y[q] = (R *)d->peer_y_final[q]; p.peer_t[q] = (R *)d->peer_t_final[q]; }
  p.keys = d->bm_keys;
  p.traj_args = (const R *)d->traj_args; p.n_traj_args = d->n_traj_args;
  p.reject_ts = nullptr; p.n_reject = d->store_rejected_steps > 0 ? d->store_rejected_steps : 0;
  p.state_in = (const R *)d->state_in; p.state_out = (R *)d->state_out; p.state_in_flags = d->state_in_flags;
  p.n_events = d->n_events; p.event_root = d->event_root_find;
  p.ev_rtol = (R)d->event_rtol; p.ev_atol = (R)d->event_atol;
  {
    int off = 0;
    for (int i = 0; i < DFX_MAX_EVENTS; ++i) {
      p.event_kind[i] = DFX_EVENT_NONE; p.event_dir[i] = 0;
      for (int c = 0; c < 4; ++c) p.ev_w[i][c] = R(0);
      p.ev_b[i] = p.ev_wt[i] = p.ev_ss_rtol[i] = p.ev_ss_atol[i] = R(0);
      p.ev_user[i] = 0;
      if (i >= d->n_events) continue;
      p.event_kind[i] = d->event_kind[i]; p.event_dir[i] = d->event_direction[i];
      if (d->event_kind[i] == DFX_EVENT_AFFINE) {
        for (int c = 0; c < d->dim && c < 4; ++c) p.ev_w[i][c] = (R)d->event_params[off + c];
        p.ev_b[i] = (R)d->event_params[off + d->dim]; p.ev_wt[i] = (R)d->event_params[off + d->dim + 1];
        off += d->dim + 2;
      } else if (d->event_kind[i] == DFX_EVENT_STEADY_STATE) {
        p.ev_ss_rtol[i] = (R)d->event_params[off]; p.ev_ss_atol[i] = (R)d->event_params[off + 1];
        off += 2;
      } else if (d->event_kind[i] == DFX_EVENT_USER) {
        p.ev_user[i] = (int)d->event_params[off];
        off += 1;
      }
    }
  }
  if (sde) {
    p.vbt.t0 = d->bm_t0; p.vbt.t1 = d->bm_t1;
    const double tol_n = d->bm_tol / (d->bm_t1 - d->bm_t0);  // tree.py:286
    int depth = 0;
    while (std::ldexp(1.0, -depth) > tol_n && depth < 1000) ++depth;  // tree.py:412
    p.vbt.depth = depth;
    p.vbt.levy = d->levy_area;
    p.vbt.partitionable = d->threefry_partitionable;
  }
  return 0;
}

struct DeviceInfo { int sms; };
constexpr long long kMoreBlocksMaxGenerations = 1;  // see launch_solve: when the one-more-CTA instantiation is preferred
int device_sm_count(int *sms);  // cached per device (api.cu)

// resident lanes of the persistent grid of one instantiation (no dynamic shared memory: the SaveAt(t1) ODE kernels)
template <class R, class Field, class Solver, int LEVY, bool RICH, bool EXTRA, bool SPEC, int MOREB>
long long resident_lanes(int sms) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ensemble_kernel<R, Field, Solver, LEVY, RICH, EXTRA, SPEC, MOREB>,
                                                    kBlockThreads, 0) != cudaSuccess) return 0;
  return (long long)sms * per_sm * kBlockThreads;
}

template <class R, class Field, class Solver, int LEVY, bool RICH, bool EXTRA = false, bool SPEC = false, int MOREB = 0>
int launch_variant(SolveParams<R> &p, const typename Field::template P<R> &fp, cudaStream_t stream) {
  auto kern = ensemble_kernel<R, Field, Solver, LEVY, RICH, EXTRA, SPEC, MOREB>;
  int sms = 0;
  if (int rc = device_sm_count(&sms)) return rc;
  // SaveAt(dense=True): per-lane staging records for the warp-cooperative stores
  size_t smem = 0;
  p.dense_coop = 0;
  {
    // dense records with write-back stores: consecutive steps of a trajectory are adjacent in every dense array (32 B of
    // y0 / y1, 8 B of ts, 448 B of k per step for Dopri8, d = 4), and the ~38 k resident trajectories' open lines fit the
    // L2 many times over, so L2 merges them into full lines before they go to HBM; evict-first (st.global.cs) stores send
    // the partial sectors out one by one (C3: 27.1 ms vs 25.7 ms).  DFX_DENSE_CS=1 selects the streaming stores.
    static const int env_cs = [] { const char *e = getenv("DFX_DENSE_CS"); return e ? atoi(e) : 0; }();
    p.dense_cs = env_cs;
    // measured on C3 (ms): scalar pad + scalar flush 26.2, scalar pad + 32-byte flush 24.6, 32-byte pad + scalar flush 28.0,
    // both 32-byte 26.9 - the record flush gains from moving a record with half a warp, while 1 KB-per-instruction padding
    // bursts from the finalising warp crowd out the other warps' record stores
    static const int env_pv = [] { const char *e = getenv("DFX_PAD_VEC"); return e ? atoi(e) : 0; }();
    static const int env_fv = [] { const char *e = getenv("DFX_FLUSH_VEC"); return e ? atoi(e) : 1; }();
    p.pad_vec = env_pv; p.flush_vec = env_fv;
  }
  {
    // finalize/refill batching (ensemble_kernel.cuh): 2 is within a few percent of the optimum sqrt(2048 X / (n I)) for
    // anything from short to long trajectories; DFX_REFILL_BATCH overrides it for experiments
    static const int env_batch = [] { const char *e = getenv("DFX_REFILL_BATCH"); return e ? atoi(e) : 0; }();
    p.refill_batch = env_batch >= 1 ? (env_batch > 32 ? 32 : env_batch) : 2;
  }
  // SDE kernels: the VBT descent cache (vbt.cuh), as many tree levels as fit ~24 KB per CTA (all of them for C5)
  p.dense_smem_offset = 0;
  p.vbt.cache_levels = 0;
  p.vbt.cache_stride = kBlockThreads;
  if (LEVY != DFX_LEVY_NONE) {
    static const int env_cache = [] { const char *e = getenv("DFX_VBT_CACHE"); return e ? atoi(e) : 1; }();
    constexpr int kWordsPerLevel = VbtCache<R, LEVY == DFX_LEVY_SPACE_TIME, NoiseDim<Field>::value>::kWords;
    const size_t per_level = (size_t)kWordsPerLevel * kBlockThreads * sizeof(R);
    constexpr int kTrees = 1;  // one tree per trajectory, whatever the Brownian shape
    int levels = env_cache ? (int)((24 * 1024) / (per_level * kTrees)) : 0;
    if (levels > p.vbt.depth) levels = p.vbt.depth;
    if (levels > 32) levels = 32;
    if (levels >= 2) {
      p.vbt.cache_levels = levels;
      smem = ((levels * per_level * kTrees + 15) / 16) * 16;
      p.dense_smem_offset = (int)smem;
    }
  }
  if (RICH && p.save_dense && (Solver::kInterp == kInterpLinear || p.dense_k != nullptr)) {
    const int kk = Solver::kInterp != kInterpLinear ? Solver::S * Field::kDim : 0;
    // must match kDenseStride in ensemble_kernel.cuh
    const int rec_n = kk + 2 * Field::kDim, vw = 32 / (int)sizeof(R);
    const bool vec = (Field::kDim % vw == 0) && (kk % vw == 0) && (rec_n / vw <= 16);
    const int stride = vec ? (((rec_n * (int)sizeof(R) + 15) / 16) | 1) * 16 / (int)sizeof(R) : (rec_n | 1);
    const size_t rec = (size_t)kBlockThreads * stride * sizeof(R);
    static const int env_coop = [] { const char *e = getenv("DFX_DENSE_COOP"); return e ? atoi(e) : 1; }();
    if (smem + rec <= 100 * 1024 && env_coop) {
      p.dense_coop = 1;
      smem += rec;
    }
  }
  if (smem > 48 * 1024) DFX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  DFX_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBlockThreads, smem));
  if (per_sm < 1) per_sm = 1;
  // persistent grid: a whole number of CTAs per SM, but never more threads than trajectories
  long long blocks = (long long)sms * per_sm;
  const long long need = (p.n_traj + kBlockThreads - 1) / kBlockThreads;
  if (blocks > need) blocks = need < 1 ? 1 : need;
  kern<<<(unsigned)blocks, kBlockThreads, smem, stream>>>(p, fp);
  count_launch();
  DFX_CUDA_OK(cudaGetLastError());
  return 0;
}


// The launcher bound into the registry for one (R, Field, Solver, LEVY) combination.
template <class R, class Field, class Solver, int LEVY>
int launch_solve(const dfx_solve_desc *d, void *stream_v) {
  cudaStream_t stream = (cudaStream_t)stream_v;
  constexpr bool SDE = LEVY != DFX_LEVY_NONE;
  SolveParams<R> p;
  fill_params<R, Solver>(d, p, SDE);
  if (d->n_field_params < Field::kNumParams) { set_error("field needs %d parameters, got %d", Field::kNumParams, d->n_field_params); return DFX_ERR_BAD_ARGUMENT; }
  const auto fp = Field::template make<R>(d->field_params, d->n_field_params, d->field_weights);
  if (p.n_traj == 0) return 0;
  if (d->traj_args != nullptr && (!PerTrajArgs<Field>::value || d->n_traj_args != Field::kNumParams)) {
    set_error("per-trajectory args: this functor takes %d (got %d)", PerTrajArgs<Field>::value ? Field::kNumParams : 0, d->n_traj_args);
    return DFX_ERR_BAD_ARGUMENT;
  }
  for (int i = 0; i < d->n_events && i < DFX_MAX_EVENTS; ++i)
    if (d->event_kind[i] == DFX_EVENT_USER && (p.ev_user[i] < 0 || p.ev_user[i] >= UserEvents<Field>::value)) {
      set_error("event %d asks for condition %d of a functor that defines %d", i, p.ev_user[i], UserEvents<Field>::value);
      return DFX_ERR_BAD_ARGUMENT;
    }
  if constexpr (SDE) {
    // the Brownian motion's shape is compiled into the functor: shape () / (1,) -> 1 component, (m,) -> m
    if ((d->bm_dim == 0 ? 1 : d->bm_dim) != NoiseDim<Field>::value) {
      if (d->bm_dim == 0) set_error("this functor is driven by a Brownian motion with %d components, got VirtualBrownianTree shape ()", NoiseDim<Field>::value);
      else set_error("this functor is driven by a Brownian motion with %d component(s), got VirtualBrownianTree shape (%d,)", NoiseDim<Field>::value, d->bm_dim);
      return DFX_ERR_BAD_ARGUMENT;
    }
    if (StateNoise<Field>::value && InnerId<Solver>::value == DFX_SHARK) {
      set_error("ShARK is an additive-noise SRK (shark.py:10-30): the diffusion of this field depends on y");
      return DFX_ERR_BAD_ARGUMENT;
    }
  }
  // EXTRA: ClipStepSizeController / Hairer starting step / Event; RICH: any SaveAt mode beyond t1 (EXTRA implies RICH)
  const bool extra = (d->hairer_initial_step && std::isnan(d->dt0)) || d->step_ts || d->jump_ts || d->n_events != 0 ||
                     d->state_in || d->state_out || d->store_rejected_steps > 0;
  const bool rich = extra || d->save_t0 || d->save_ts || d->save_steps || d->save_dense;
  if (HostPipe *hp = host_pipe()) {
    if (rich) { set_error("internal: the host pipeline only drives the SaveAt(t1=True) kernel"); return DFX_ERR_BAD_ARGUMENT; }
    p.pipe_in_ready = hp->in_ready;
    p.pipe_done = hp->done;
    p.pipe_host_flags = hp->host_flags;
    p.pipe_chunk_len = hp->chunk_len;
    hp->consumed = true;
  }

  // scratch: the work-queue counter.  (The +inf padding of unfilled output slots is written by the solve kernel itself
  // when it finalises a trajectory, so there is no second pass over the buffers.)
  unsigned long long *counter = nullptr;
  char *scratch = nullptr;
  const size_t reject_bytes = p.n_reject > 0 ? (size_t)p.n_traj * p.n_reject * sizeof(R) : 0;  // rejected-times stacks
  DFX_CUDA_OK(cudaMallocAsync((void **)&scratch, 16 + reject_bytes, stream));
  DFX_CUDA_OK(cudaMemsetAsync(scratch, 0, 16, stream));
  counter = (unsigned long long *)scratch;
  p.work_counter = counter;
  if (reject_bytes) p.reject_ts = (R *)(scratch + 16);

  if (p.totals) DFX_CUDA_OK(cudaMemsetAsync(p.totals, 0, 4 * sizeof(long long), stream));
  int rc;
  // fp64 ODE solves with the default controller configuration run the specialised instantiations (see SPEC)
  constexpr bool kHasSpec = DFX_OPT_FAST_PID && IsTableau<Solver>::value && !IsHalf<Solver>::value && !SDE && sizeof(R) == 8;
  bool spec = false;
  if constexpr (kHasSpec) spec = p.controller == DFX_CTRL_PID && p.fast_pid && !p.has_dtmin && !p.has_dtmax;
  if (extra) rc = launch_variant<R, Field, Solver, LEVY, true, true>(p, fp, stream);
  else if (rich) {
    if constexpr (kHasSpec) {
      if (spec) rc = launch_variant<R, Field, Solver, LEVY, true, false, true>(p, fp, stream);
      else rc = launch_variant<R, Field, Solver, LEVY, true>(p, fp, stream);
    } else {
      rc = launch_variant<R, Field, Solver, LEVY, true>(p, fp, stream);
    }
  } else {
    if constexpr (kHasSpec) {
      if (spec) {
        // one resident generation if one more CTA per SM makes the whole batch fit (see MOREB)
        int sms = 0;
        if (int e = device_sm_count(&sms)) return e;
        static const int env_moreb = [] { const char *e = getenv("DFX_MORE_BLOCKS"); return e ? atoi(e) : 1; }();
        const long long base = resident_lanes<R, Field, Solver, LEVY, false, false, true, 0>(sms);
        const long long more = env_moreb ? resident_lanes<R, Field, Solver, LEVY, false, false, true, 1>(sms) : 0;
        // fewer (and fuller) generations at the higher occupancy; with many generations the queue evens the tail out by
        // itself and the default occupancy has the better per-warp throughput (C2 at 2^20: 2.97 vs 3.05 ms)
        const bool fewer = base > 0 && more > base && p.n_traj > base &&
                           (p.n_traj + more - 1) / more < (p.n_traj + base - 1) / base && (p.n_traj + more - 1) / more <= kMoreBlocksMaxGenerations;
        if (env_moreb == 2 || (env_moreb == 1 && fewer)) rc = launch_variant<R, Field, Solver, LEVY, false, false, true, 1>(p, fp, stream);
        else rc = launch_variant<R, Field, Solver, LEVY, false, false, true>(p, fp, stream);
      }
      else rc = launch_variant<R, Field, Solver, LEVY, false>(p, fp, stream);
    } else {
      rc = launch_variant<R, Field, Solver, LEVY, false>(p, fp, stream);
    }
  }
  cudaFreeAsync(scratch, stream);
  return rc;
}

template <class R> struct DtypeOf;
template <> struct DtypeOf<double> { static constexpr int value = DFX_F64; };
template <> struct DtypeOf<float> { static constexpr int value = DFX_F32; };

template <class R, class Field, class Solver, int LEVY>
struct Registrar {
  Registrar() { register_builtin(Field::kId, Field::kDim, Solver::kId, DtypeOf<R>::value, LEVY, &launch_solve<R, Field, Solver, LEVY>); }
};

#define DFX_CAT2(a, b) a##b
#define DFX_CAT(a, b) DFX_CAT2(a, b)
#define DFX_REGISTER(R, Field, Solver, LEVY) \
  static ::dfx::Registrar<R, Field, Solver, LEVY> DFX_CAT(dfx_registrar_, __COUNTER__);

// all explicit RK tableaux for an ODE field, both dtypes
#define DFX_REGISTER_ODE_FIELD(Field)                     \
  DFX_REGISTER(double, Field, ::dfx::Tsit5, 0)            \
  DFX_REGISTER(double, Field, ::dfx::Dopri5, 0)           \
  DFX_REGISTER(double, Field, ::dfx::Dopri8, 0)           \
  DFX_REGISTER(double, Field, ::dfx::Heun, 0)             \
  DFX_REGISTER(double, Field, ::dfx::Bosh3, 0)            \
  DFX_REGISTER(double, Field, ::dfx::Midpoint, 0)         \
  DFX_REGISTER(double, Field, ::dfx::Ralston, 0)          \
  DFX_REGISTER(double, Field, ::dfx::EulerSolver, 0)      \
  DFX_REGISTER(float, Field, ::dfx::Tsit5, 0)             \
  DFX_REGISTER(float, Field, ::dfx::Dopri5, 0)            \
  DFX_REGISTER(float, Field, ::dfx::Dopri8, 0)            \
  DFX_REGISTER(float, Field, ::dfx::Heun, 0)              \
  DFX_REGISTER(float, Field, ::dfx::Bosh3, 0)             \
  DFX_REGISTER(float, Field, ::dfx::Midpoint, 0)          \
  DFX_REGISTER(float, Field, ::dfx::Ralston, 0)           \
  DFX_REGISTER(float, Field, ::dfx::EulerSolver, 0) \
  DFX_REGISTER(double, Field, ::dfx::HalfOf<::dfx::EulerSolver>, 0) \
  DFX_REGISTER(double, Field, ::dfx::HalfOf<::dfx::Heun>, 0)        \
  DFX_REGISTER(float, Field, ::dfx::HalfOf<::dfx::EulerSolver>, 0)  \
  DFX_REGISTER(float, Field, ::dfx::HalfOf<::dfx::Heun>, 0)

}  // namespace dfx
