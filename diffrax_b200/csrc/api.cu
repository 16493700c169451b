// api.cu - C ABI of libdiffrax_b200 (include/diffrax_b200.h): registry, argument checking,
// host-buffer wrapper, and the small standalone kernels (PRNG known-answer entry points,
// VirtualBrownianTree.evaluate, DenseInterpolation.evaluate, pipe-peak microbenchmarks).
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "launch.cuh"

namespace dfx {

static thread_local char tl_error[512] = "";
static thread_local long long tl_launches = 0;

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { tl_launches += n; }
static thread_local HostPipe *tl_host_pipe = nullptr;
HostPipe *&host_pipe() { return tl_host_pipe; }

using RegKey = std::tuple<int, int, int, int, int>;  // field, dim, solver, dtype, levy
static std::map<RegKey, dfx_launcher_fn> &registry() {
  static std::map<RegKey, dfx_launcher_fn> r;
  return r;
}
static std::mutex &registry_mutex() {
  static std::mutex m;
  return m;
}
int register_builtin(int field_id, int dim, int solver_id, int dtype, int levy, dfx_launcher_fn fn) {
  std::lock_guard<std::mutex> g(registry_mutex());
  registry()[RegKey(field_id, dim, solver_id, dtype, levy)] = fn;
  return 0;
}
static dfx_launcher_fn find_launcher(int field_id, int dim, int solver_id, int dtype, int levy) {
  std::lock_guard<std::mutex> g(registry_mutex());
  auto it = registry().find(RegKey(field_id, dim, solver_id, dtype, levy));
  return it == registry().end() ? nullptr : it->second;
}

int device_sm_count(int *sms) {
  static std::mutex m;
  static std::map<int, int> cache;
  int dev = 0;
  DFX_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> g(m);
  auto it = cache.find(dev);
  if (it == cache.end()) {
    int n = 0;
    DFX_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    // keep stream-ordered scratch cached in the default pool instead of returning it at every sync
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    it = cache.emplace(dev, n).first;
  }
  *sms = it->second;
  return 0;
}

// ------------------------------------------------------------------------------------------
// standalone kernels
// ------------------------------------------------------------------------------------------
__global__ void threefry_kernel(long long n, const uint32_t *keys, const uint32_t *ctrs, uint32_t *out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t a, b;
  threefry2x32(keys[2 * i], keys[2 * i + 1], ctrs[2 * i], ctrs[2 * i + 1], a, b);
  out[2 * i] = a;
  out[2 * i + 1] = b;
}

__global__ void split_kernel(long long n, const uint32_t *keys, int num, int partitionable, uint32_t *out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t k0 = keys[2 * i], k1 = keys[2 * i + 1];
  uint32_t *o = out + i * 2 * num;
  for (int c = 0; c < num; ++c) {
    uint32_t a, b;
    if (partitionable) {
      threefry2x32(k0, k1, 0u, (uint32_t)c, a, b);
      o[2 * c] = a;
      o[2 * c + 1] = b;
    } else {
      threefry2x32(k0, k1, (uint32_t)c, (uint32_t)(num + c), a, b);
      o[c] = a;
      o[num + c] = b;
    }
  }
}

// dst[0..n) = *src (a device-resident scalar; the jax.ffi shim's unbatched traced t0 / t1)
template <class R> __global__ void broadcast_kernel(long long n, const R *src, R *dst) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = *src;
}

// jax.random.normal(key, shape, dtype) for n keys; shape () (m == 0) or (m,)
template <class R, int M>
__global__ void normal_kernel(long long n, const uint32_t *keys, int partitionable, R *out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
#pragma unroll
  for (int w = 0; w < M; ++w) out[i * M + w] = random_normal<R, M>(Key{keys[2 * i], keys[2 * i + 1]}, partitionable != 0, w);
}

template <class R, bool STLA, int M>
__global__ void vbt_kernel(long long n, const uint32_t *keys, VbtParams vp, const R *ta, const R *tb, int per_traj,
                           R *W, R *H) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  BrownianTree<R, STLA, M> bm;
  bm.init(keys + 2 * i, vp);
  R w[M], h[M];
  bm.increment(ta[per_traj ? i : 0], tb[per_traj ? i : 0], vp, w, h);
#pragma unroll
  for (int c = 0; c < M; ++c) {
    W[i * M + c] = w[c];
    if (H) H[i * M + c] = h[c];
  }
}

template <class R, int M>
static void launch_normal(int64_t n, const uint32_t *keys, int partitionable, void *out, cudaStream_t st) {
  normal_kernel<R, M><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, keys, partitionable, (R *)out);
}
template <class R, int M>
static void launch_vbt(bool stla, int64_t n, const uint32_t *keys, const VbtParams &vp, const void *ta, const void *tb, int per_traj,
                       void *W, void *H, cudaStream_t st) {
  const unsigned blocks = (unsigned)((n + 127) / 128);
  if (stla) vbt_kernel<R, true, M><<<blocks, 128, 0, st>>>(n, keys, vp, (const R *)ta, (const R *)tb, per_traj, (R *)W, (R *)H);
  else vbt_kernel<R, false, M><<<blocks, 128, 0, st>>>(n, keys, vp, (const R *)ta, (const R *)tb, per_traj, (R *)W, (R *)H);
}
// DenseInterpolation.evaluate / .derivative (_global_interpolation.py:335-368), one thread per (trajectory, query).
template <class R, class Solver, int D, bool DERIV>
__global__ void dense_eval_kernel(long long n_traj, int max_steps, const R *dts, const R *dy0, const R *dy1,
                                  const R *dk, const int *dcount, R direction, const R *tq, int nq, R *out) {
  constexpr int S = Solver::S;
  const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (gid >= n_traj * nq) return;
  const long long i = gid / nq;
  const R *ts = dts + i * (long long)(max_steps + 1);
  const int ts_size = dcount[i] + 1;
  R t = tq[gid] * direction;
  // _nan_if_out_of_bounds (370-381)
  if (ts_size <= 1 || t < ts[0] || t > ts[ts_size - 1]) t = Num<R>::nan();
  // _interpret_t (36-45): searchsorted(ts, t, side="left") over the inf-padded array, then clip(index-1, 0, ts_size-2).
  // Searching the filled prefix [0, ts_size) gives the same index (every padding slot is +inf > t) and does not read the
  // tails, which are unwritten under dense_lazy_padding.
  int lo = 0, hi = ts_size;
  if (t != t) lo = max_steps + 1;
  else while (lo < hi) { const int mid = (lo + hi) >> 1; if (ts[mid] < t) lo = mid + 1; else hi = mid; }
  int index = lo - 1;
  if (index > ts_size - 2) index = ts_size - 2;
  if (index < 0) index = 0;
  R y0[D], y1[D], k[S][D], o[D];
  const long long row = i * (long long)max_steps + index;
#pragma unroll
  for (int c = 0; c < D; ++c) { y0[c] = dy0[row * D + c]; y1[c] = dy1[row * D + c]; }
#pragma unroll
  for (int j = 0; j < S; ++j)
#pragma unroll
    for (int c = 0; c < D; ++c) k[j][c] = (Solver::kInterp != kInterpLinear && dk) ? dk[(row * S + j) * D + c] : R(0);
  if constexpr (DERIV) {
    interp_deriv<Solver::kInterp, R, S, D>(ts[index], ts[index + 1], y0, y1, k, t, o);
#pragma unroll
    for (int c = 0; c < D; ++c) out[gid * D + c] = direction * o[c];  // 366: direction * derivative
  } else {
    interp_eval<Solver::kInterp, R, S, D>(ts[index], ts[index + 1], y0, y1, k, t, o);
#pragma unroll
    for (int c = 0; c < D; ++c) out[gid * D + c] = o[c];
  }
}

// +inf into the unfilled tails of the SaveAt(dense=True) buffers (_integrate.py:1296-1300, 1320-1322): one warp per
// trajectory, coalesced streaming stores.  Used when a solve ran with dense_lazy_padding and somebody wants the raw arrays.
template <class R>
__global__ void dense_pad_kernel(long long n_traj, int max_steps, int d, int sd, R *dts, R *dy0, R *dy1, R *dk, const int *dcount) {
  const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_traj) return;
  const long long dc = dcount[w], ms = max_steps;
  pad_tail(dts + w * (ms + 1), dc + 1, ms + 1, lane, true);
  pad_tail(dy0 + w * ms * d, dc * d, ms * d, lane, true);
  pad_tail(dy1 + w * ms * d, dc * d, ms * d, lane, true);
  if (dk) pad_tail(dk + w * ms * sd, dc * sd, ms * sd, lane, true);
}

// Pipe-peak microbenchmarks: 8 independent FMA chains per thread, enough warps to fill every SM.
template <class R>
__global__ void fma_peak_kernel(R *out, int iters, R a, R b) {
  R x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = (R)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = x[i] * a + b;
  }
  R s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == (R)123456789) out[0] = s;
}
__global__ void int_peak_kernel(uint32_t *out, int iters) {
  uint32_t x[8], y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x + i; y[i] = blockIdx.x * 7 + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] += y[i]; y[i] = rotl(y[i], 13); y[i] ^= x[i]; }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] ^ y[i];
  if (s == 0x12345678u) out[0] = s;
}

template <class R, class Solver>
static int dense_eval_dispatch_dim(bool deriv, int dim, long long n, int ms, const void *dts, const void *dy0, const void *dy1,
                                   const void *dk, const int *dc, double direction, const void *tq, int nq, void *out,
                                   cudaStream_t st) {
  const long long total = n * nq;
  const unsigned blocks = (unsigned)((total + 127) / 128);
#define DFX_DE(DD)                                                                                              \
  case DD:                                                                                                      \
    if (deriv)                                                                                                  \
      dense_eval_kernel<R, Solver, DD, true><<<blocks, 128, 0, st>>>(n, ms, (const R *)dts, (const R *)dy0, (const R *)dy1, \
                                                                     (const R *)dk, dc, (R)direction, (const R *)tq, nq, (R *)out); \
    else                                                                                                        \
      dense_eval_kernel<R, Solver, DD, false><<<blocks, 128, 0, st>>>(n, ms, (const R *)dts, (const R *)dy0, (const R *)dy1, \
                                                                      (const R *)dk, dc, (R)direction, (const R *)tq, nq, (R *)out); \
    break;
  switch (dim) {
    DFX_DE(1) DFX_DE(2) DFX_DE(3) DFX_DE(4)
    default:
      set_error("dense_evaluate: unsupported dim %d", dim);
      return DFX_ERR_UNSUPPORTED;
  }
#undef DFX_DE
  count_launch();
  DFX_CUDA_OK(cudaGetLastError());
  return 0;
}

template <class R>
static int dense_eval_dispatch(bool deriv, int solver_id, int dim, long long n, int ms, const void *dts, const void *dy0,
                               const void *dy1, const void *dk, const int *dc, double direction, const void *tq, int nq,
                               void *out, cudaStream_t st) {
  switch (solver_id & ~DFX_HALF_SOLVER) {  // HalfSolver(inner) records the inner solver's dense_info
#define DFX_DS(ID, T) case ID: return dense_eval_dispatch_dim<R, T>(deriv, dim, n, ms, dts, dy0, dy1, dk, dc, direction, tq, nq, out, st);
    DFX_DS(DFX_TSIT5, Tsit5) DFX_DS(DFX_DOPRI5, Dopri5) DFX_DS(DFX_DOPRI8, Dopri8) DFX_DS(DFX_HEUN, Heun)
    DFX_DS(DFX_BOSH3, Bosh3) DFX_DS(DFX_MIDPOINT, Midpoint) DFX_DS(DFX_RALSTON, Ralston)
    DFX_DS(DFX_EULER, EulerSolver) DFX_DS(DFX_SHARK, SharkSolver)
#undef DFX_DS
  }
  set_error("dense_evaluate: unknown solver %d", solver_id);
  return DFX_ERR_BAD_ARGUMENT;
}

}  // namespace dfx

using namespace dfx;

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int dfx_abi_version(void) { return DFX_ABI_VERSION; }
const char *dfx_last_error(void) { return tl_error; }
int64_t dfx_launch_count(void) { return tl_launches; }
void dfx_reset_launch_count(void) { tl_launches = 0; }

int dfx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int dfx_num_stages(int solver_id) {
  switch (solver_id & ~DFX_HALF_SOLVER) {  // HalfSolver(inner): the inner solver's stages (its dense_info)
    case DFX_TSIT5: return Tsit5::S; case DFX_DOPRI5: return Dopri5::S; case DFX_DOPRI8: return Dopri8::S;
    case DFX_HEUN: return Heun::S; case DFX_BOSH3: return Bosh3::S; case DFX_MIDPOINT: return Midpoint::S;
    case DFX_RALSTON: return Ralston::S; case DFX_EULER: return 1; case DFX_SHARK: return 2;
  }
  return -1;
}
int dfx_solver_order(int solver_id) {
  switch (solver_id & ~DFX_HALF_SOLVER) {
    case DFX_TSIT5: return Tsit5::kOrder; case DFX_DOPRI5: return Dopri5::kOrder; case DFX_DOPRI8: return Dopri8::kOrder;
    case DFX_HEUN: return Heun::kOrder; case DFX_BOSH3: return Bosh3::kOrder; case DFX_MIDPOINT: return Midpoint::kOrder;
    case DFX_RALSTON: return Ralston::kOrder; case DFX_EULER: return 1; case DFX_SHARK: return 2;
  }
  return -1;
}
int dfx_field_dim(int field_id) {
  switch (field_id) {
    case DFX_FIELD_DECAY: return 0;
    case DFX_FIELD_LOTKA_VOLTERRA: return 2; case DFX_FIELD_LORENZ: return 3; case DFX_FIELD_CR3BP: return 4;
    case DFX_FIELD_OU: return 0; case DFX_FIELD_FORCED_OSC: return 2; case DFX_FIELD_VDP: return 2;
    case DFX_FIELD_MLP: return 0;
  }
  return -1;
}
int dfx_has_kernel(int field_id, int dim, int solver_id, int dtype, int levy_area) {
  return find_launcher(field_id, dim, solver_id, dtype, levy_area) != nullptr;
}
int dfx_register_launcher(int field_id, int dim, int solver_id, int dtype, int levy_area, dfx_launcher_fn fn) {
  if (!fn) { set_error("null launcher"); return DFX_ERR_BAD_ARGUMENT; }
  return register_builtin(field_id, dim, solver_id, dtype, levy_area, fn);
}

// _integrate.py:1273-1293
int dfx_out_size(const dfx_solve_desc *d) {
  int out = 0;
  if (d->save_t0) out += 1;
  if (d->save_ts) out += d->n_save_ts;
  if (d->save_steps != 0) out += d->max_steps / d->save_steps;
  if (d->save_t1 && (d->save_steps == 0 || (d->max_steps % d->save_steps) != 0)) out += 1;
  return out;
}

static int check_desc(const dfx_solve_desc *d) {
  if (!d) { set_error("null descriptor"); return DFX_ERR_BAD_ARGUMENT; }
  if (d->struct_size != sizeof(dfx_solve_desc) || d->abi_version != DFX_ABI_VERSION) {
    set_error("descriptor ABI mismatch: size %u (want %zu), version %u (want %d)", d->struct_size,
              sizeof(dfx_solve_desc), d->abi_version, DFX_ABI_VERSION);
    return DFX_ERR_BAD_ARGUMENT;
  }
  // (one thread per trajectory: dim <= 8; user functors may be "wide" - one warp per trajectory, csrc/wide_kernel.cuh - up to 1024)
  const int max_dim = d->field_id >= DFX_FIELD_USER ? 1024 : kMaxDim;
  if (d->n_traj < 0 || d->n_traj > 0x7fffffffLL || d->dim < 1 || d->dim > max_dim) { set_error("bad n_traj (0 .. 2^31-1 per call) / dim (1 .. %d)", max_dim); return DFX_ERR_BAD_ARGUMENT; }
  if (d->dtype != DFX_F64 && d->dtype != DFX_F32) { set_error("bad dtype %d", d->dtype); return DFX_ERR_BAD_ARGUMENT; }
  if (d->n_traj > 0 && !d->y0) { set_error("y0 is null"); return DFX_ERR_BAD_ARGUMENT; }
  if (d->n_traj > 0 && (!d->stats || !d->result)) { set_error("stats / result buffers are required"); return DFX_ERR_BAD_ARGUMENT; }
  if (d->max_steps < 0) { set_error("max_steps must be >= 0 (max_steps=None is not supported)"); return DFX_ERR_BAD_ARGUMENT; }
  if (d->traj_args != nullptr && (d->n_traj_args < 1 || d->n_traj_args > 64)) { set_error("traj_args needs 1 <= n_traj_args <= 64"); return DFX_ERR_BAD_ARGUMENT; }
  if (d->save_steps < 0) { set_error("save_steps must be >= 0"); return DFX_ERR_BAD_ARGUMENT; }
  if (d->controller == DFX_CTRL_CONSTANT && is_nan(d->dt0)) {
    // constant.py:41-45
    set_error("Constant step size solvers cannot select step size automatically; please pass a value for `dt0`.");
    return DFX_ERR_BAD_ARGUMENT;
  }
  if ((d->step_ts || d->jump_ts) && d->controller != DFX_CTRL_PID) {
    // clip.py:203-207
    set_error("Can only apply `ClipStepSizeController` to adaptive step size controllers.");
    return DFX_ERR_BAD_ARGUMENT;
  }
  if (d->controller == DFX_CTRL_PID && d->levy_area != DFX_LEVY_NONE && (d->solver_id & ~DFX_HALF_SOLVER) == DFX_EULER) {
    // _integrate.py:1143-1149 ("Specific check to not work even if using HalfSolver(Euler())")
    set_error("An SDE should not be solved with adaptive step sizes with Euler's method, as it may not converge to the correct solution.");
    return DFX_ERR_BAD_ARGUMENT;
  }
  if (d->controller == DFX_CTRL_PID && d->solver_id == DFX_EULER) {
    // pid.py:461-469 (Euler provides no error estimate)
    set_error("Cannot use adaptive step sizes with a solver that does not provide error estimates.");
    return DFX_ERR_BAD_ARGUMENT;
  }
  if (!d->t0_per_traj && !d->t1_per_traj && !is_nan(d->dt0) && (d->t1 - d->t0) * d->dt0 < 0) {
    set_error("Must have (t1 - t0) * dt0 >= 0");  // _integrate.py:1036-1045
    return DFX_ERR_BAD_ARGUMENT;
  }
  if (d->n_events != 0) {
    int need = 0;
    bool ok = d->n_events > 0 && d->n_events <= DFX_MAX_EVENTS && d->event_params != nullptr;
    for (int i = 0; ok && i < d->n_events; ++i) {
      if (d->event_kind[i] == DFX_EVENT_AFFINE) { need += d->dim + 2; ok = d->dim <= 4; }
      else if (d->event_kind[i] == DFX_EVENT_STEADY_STATE) need += 2;
      else if (d->event_kind[i] == DFX_EVENT_USER) { need += 1; ok = d->field_id >= DFX_FIELD_USER; }
      else ok = false;
      if (d->event_direction[i] < 0 || d->event_direction[i] > 2) ok = false;
    }
    if (!ok || d->n_event_params != need) {
      set_error("bad event description: 1..%d events of kind affine (dim + 2 params, dim <= 4) / steady state (2 params) / user "
                "(1 param, user functors only), direction in {0,1,2}; got %d events, %d params (need %d)", DFX_MAX_EVENTS, d->n_events, d->n_event_params, need);
      return DFX_ERR_BAD_ARGUMENT;
    }
  }
  if (d->n_peers < 0 || d->n_peers > DFX_MAX_PEERS) { set_error("n_peers must be 0 .. %d", DFX_MAX_PEERS); return DFX_ERR_BAD_ARGUMENT; }
  for (int q = 0; q < d->n_peers; ++q)
    if (!d->peer_y_final[q] || !d->peer_t_final[q]) { set_error("peer buffer %d is null", q); return DFX_ERR_BAD_ARGUMENT; }
  const int T = dfx_out_size(d);
  if (d->n_traj > 0 && T > 0 && (!d->ts_out || !d->ys_out)) { set_error("ts_out / ys_out are required (T_out = %d)", T); return DFX_ERR_BAD_ARGUMENT; }
  if (d->n_traj > 0 && d->save_dense && (!d->dense_ts || !d->dense_y0 || !d->dense_y1 || !d->dense_count)) {
    set_error("dense buffers are required for SaveAt(dense=True)");
    return DFX_ERR_BAD_ARGUMENT;
  }
  if (d->levy_area != DFX_LEVY_NONE) {
    if (!d->bm_keys) { set_error("SDE solve needs bm_keys"); return DFX_ERR_BAD_ARGUMENT; }
    const bool matrix = d->field_id > DFX_FIELD_OU_MATRIX && d->field_id <= DFX_FIELD_OU_MATRIX + 4;
    if (matrix && d->bm_dim != d->field_id - DFX_FIELD_OU_MATRIX) {
      set_error("a [d, %d] diffusion matrix needs VirtualBrownianTree(shape=(%d,)), got shape (%d,)", d->field_id - DFX_FIELD_OU_MATRIX,
                d->field_id - DFX_FIELD_OU_MATRIX, d->bm_dim);
      return DFX_ERR_BAD_ARGUMENT;
    }
    // (user functors, field ids >= DFX_FIELD_USER, define their own noise shape: the launcher checks bm_dim against it)
    if (!matrix && d->field_id < DFX_FIELD_USER && d->bm_dim != 0 && d->bm_dim != d->dim) {
      set_error("VirtualBrownianTree(shape=(%d,)) drives a diagonal diffusion: it needs a state of the same dimension (got %d)", d->bm_dim, d->dim);
      return DFX_ERR_BAD_ARGUMENT;
    }
    if (d->bm_dim == 0 && d->dim != 1 && d->field_id == DFX_FIELD_OU) {
      set_error("the OU functor with a %d-dimensional state needs VirtualBrownianTree(shape=(%d,))", d->dim, d->dim);
      return DFX_ERR_BAD_ARGUMENT;
    }
    if (d->field_id == DFX_FIELD_GBM && (d->solver_id & ~DFX_HALF_SOLVER) == DFX_SHARK) {
      set_error("ShARK is an additive-noise SRK (shark.py:10-30): the diffusion of this field depends on y");
      return DFX_ERR_BAD_ARGUMENT;
    }
    if (!(d->bm_t0 < d->bm_t1)) { set_error("t0 must be strictly less than t1"); return DFX_ERR_BAD_ARGUMENT; }  // tree.py:281
    if ((d->solver_id & ~DFX_HALF_SOLVER) == DFX_SHARK && d->levy_area != DFX_LEVY_SPACE_TIME) {
      set_error("The Brownian increment does not have the minimal Levy Area SpaceTimeLevyArea.");  // srk.py:391-395
      return DFX_ERR_BAD_ARGUMENT;
    }
  } else if ((d->solver_id & ~DFX_HALF_SOLVER) == DFX_SHARK) {
    set_error("ShARK needs MultiTerm(ODETerm, ControlTerm(VirtualBrownianTree))");
    return DFX_ERR_BAD_ARGUMENT;
  }
  return 0;
}

int dfx_ensemble_solve(const dfx_solve_desc *d, void *cuda_stream) {
  if (int rc = check_desc(d)) return rc;
  if (d->n_traj == 0) {  // an empty batch: nothing to integrate (the totals, if wanted, are zero)
    if (d->totals) DFX_CUDA_OK(cudaMemsetAsync(d->totals, 0, 4 * sizeof(int64_t), (cudaStream_t)cuda_stream));
    return 0;
  }
  dfx_launcher_fn fn = find_launcher(d->field_id, d->dim, d->solver_id, d->dtype, d->levy_area);
  if (!fn) {
    set_error("no kernel registered for field %d dim %d solver %d dtype %d levy %d", d->field_id, d->dim,
              d->solver_id, d->dtype, d->levy_area);
    return DFX_ERR_UNSUPPORTED;
  }
  return fn(d, cuda_stream);
}

// Host-buffer variant.  The batch is cut into chunks that are pipelined over two streams (H2D of chunk i+1 and D2H of
// chunk i-1 overlap the kernel of chunk i); trajectories are independent, so chunking does not change any result.
// Pass pinned host pointers for full PCIe speed.  Dense output is not chunked (its buffers are large; one chunk).
static int solve_host_chunk(const dfx_solve_desc *h, int64_t lo, int64_t cnt, cudaStream_t st, std::vector<void *> &allocs,
                            int64_t *chunk_totals /* host [4] or null */) {
  const size_t es = h->dtype == DFX_F64 ? 8 : 4;
  const size_t N = (size_t)cnt, D = (size_t)h->dim;
  const int T = dfx_out_size(h);
  const int S = dfx_num_stages(h->solver_id);
  const size_t ms = (size_t)h->max_steps;
  dfx_solve_desc d = *h;
  d.n_traj = cnt;
  std::vector<std::tuple<void *, void *, size_t>> d2h;  // host dst, device src, bytes
  int rc = 0;
  auto off = [&](const void *base, size_t per_traj_bytes) -> const char * {
    return base ? (const char *)base + (size_t)lo * per_traj_bytes : nullptr;
  };
  auto dev_in = [&](const void *src, size_t bytes) -> void * {
    if (!src || bytes == 0 || rc) return nullptr;
    void *p = nullptr;
    if (cudaMallocAsync(&p, bytes, st) != cudaSuccess) { set_error("cudaMallocAsync(%zu) failed", bytes); rc = DFX_ERR_CUDA; return nullptr; }
    allocs.push_back(p);
    if (cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) { set_error("H2D copy failed"); rc = DFX_ERR_CUDA; }
    return p;
  };
  auto dev_out = [&](const void *dst, size_t bytes) -> void * {
    if (!dst || bytes == 0 || rc) return nullptr;
    void *p = nullptr;
    if (cudaMallocAsync(&p, bytes, st) != cudaSuccess) { set_error("cudaMallocAsync(%zu) failed", bytes); rc = DFX_ERR_CUDA; return nullptr; }
    allocs.push_back(p);
    d2h.emplace_back((void *)dst, p, bytes);
    return p;
  };
  d.y0 = dev_in(off(h->y0, D * es), N * D * es);
  d.t0_per_traj = dev_in(off(h->t0_per_traj, es), N * es);
  d.t1_per_traj = dev_in(off(h->t1_per_traj, es), N * es);
  d.save_ts = dev_in(h->save_ts, (size_t)h->n_save_ts * es);
  d.step_ts = dev_in(h->step_ts, (size_t)h->n_step_ts * es);
  d.jump_ts = dev_in(h->jump_ts, (size_t)h->n_jump_ts * es);
  d.bm_keys = (const uint32_t *)dev_in(off(h->bm_keys, 8), N * 8);
  d.field_weights = dev_in(h->field_weights, (size_t)h->n_field_weights * es);
  d.traj_args = dev_in(off(h->traj_args, (size_t)h->n_traj_args * es), N * (size_t)h->n_traj_args * es);
  d.ts_out = dev_out(off(h->ts_out, T * es), N * T * es);
  d.ys_out = dev_out(off(h->ys_out, T * D * es), N * T * D * es);
  d.stats = (int32_t *)dev_out(off(h->stats, 12), N * 3 * 4);
  d.result = (int32_t *)dev_out(off(h->result, 4), N * 4);
  if (lo == 0 && cnt == h->n_traj && h->totals_device) {
    // a single chunk: the kernel accumulates straight into the caller's device words (no host round trip)
    d.totals = h->totals_device;
    if (h->totals) d2h.emplace_back((void *)h->totals, (void *)h->totals_device, 4 * sizeof(int64_t));
  } else {
    d.totals = (int64_t *)dev_out(chunk_totals, chunk_totals ? 4 * sizeof(int64_t) : 0);  // per chunk; combined by the caller
  }
  d.totals_device = nullptr;
  d.save_count = (int32_t *)dev_out(off(h->save_count, 4), N * 4);
  if (h->save_dense) {
    d.dense_ts = dev_out(off(h->dense_ts, (ms + 1) * es), N * (ms + 1) * es);
    d.dense_y0 = dev_out(off(h->dense_y0, ms * D * es), N * ms * D * es);
    d.dense_y1 = dev_out(off(h->dense_y1, ms * D * es), N * ms * D * es);
    d.dense_k = dev_out(off(h->dense_k, ms * S * D * es), N * ms * S * D * es);
    d.dense_count = (int32_t *)dev_out(off(h->dense_count, 4), N * 4);
  }
  d.state_in = dev_in(off(h->state_in, (5 + D) * es), N * (5 + D) * es);
  d.state_out = dev_out(off(h->state_out, (5 + D) * es), N * (5 + D) * es);
  // finals: into the caller's device buffers when given (then copied to the host outputs from there), else scratch
  if (h->y_final_device) {
    d.y_final = (char *)h->y_final_device + (size_t)lo * D * es;
    if (h->y_final) d2h.emplace_back((char *)h->y_final + (size_t)lo * D * es, d.y_final, N * D * es);
  } else {
    d.y_final = dev_out(off(h->y_final, D * es), N * D * es);
  }
  if (h->t_final_device) {
    d.t_final = (char *)h->t_final_device + (size_t)lo * es;
    if (h->t_final) d2h.emplace_back((char *)h->t_final + (size_t)lo * es, d.t_final, N * es);
  } else {
    d.t_final = dev_out(off(h->t_final, es), N * es);
  }
  d.y_final_device = d.t_final_device = nullptr;
  d.peer_row_offset = h->peer_row_offset + lo;  // (the peer buffers are device pointers already: passed through)
  if (!rc) rc = dfx_ensemble_solve(&d, (void *)st);
  if (!rc)
    for (auto &c : d2h)
      if (cudaMemcpyAsync(std::get<0>(c), std::get<1>(c), std::get<2>(c), cudaMemcpyDeviceToHost, st) != cudaSuccess) {
        set_error("D2H copy failed");
        rc = DFX_ERR_CUDA;
        break;
      }
  return rc;
}

// Host-buffer variant, pipelined inside ONE launch (the SaveAt(t1=True) kernel of the built-in functors).  Cutting the
// batch into separate launches costs the work queue its depth: a 128K-trajectory chunk is ~1.15 trajectories per
// resident thread, so every chunk ends in a tail as long as its slowest trajectory.  Here the kernel runs once over the
// whole batch; the copy engines deliver the inputs chunk by chunk on their own stream, each chunk followed by a 4-byte
// copy that bumps `in_ready` (a lane that claims a trajectory of a chunk not yet resident waits on that word), and the
// kernel counts finalised trajectories per chunk and raises a flag in mapped host memory when a chunk is complete, at
// which point this thread enqueues the chunk's D2H copies.  The queue hands trajectories out in index order, so chunks
// complete roughly in order and the transfers hide behind the solve except for the first H2D and the last D2H chunk.
static constexpr int kMaxPipeChunks = 64;
static bool host_pipe_eligible(const dfx_solve_desc *h) {
  if (const char *e = std::getenv("DFX_HOST_PIPE")) { if (atoi(e) == 0) return false; }
  // the kernel waits for chunks that are enqueued after its launch returns: with synchronous launches that never happens
  if (const char *e = std::getenv("CUDA_LAUNCH_BLOCKING")) { if (atoi(e) != 0) return false; }
  const bool extra = (h->hairer_initial_step && std::isnan(h->dt0)) || h->step_ts || h->jump_ts || h->n_events != 0 ||
                     h->state_in || h->state_out || h->store_rejected_steps > 0;
  const bool rich = extra || h->save_t0 || h->save_ts || h->save_steps || h->save_dense;
  // Adaptive solves only.  Measured (B200, 2^20 trajectories): Lorenz/Dopri5/PID 4.18 ms in 8 separate launches -> 3.9 ms
  // pipelined.  A fixed-step SDE ensemble is the opposite case: equal-length trajectories leave separate launches
  // almost no tail, while inside one launch the integer-bound warps are scheduled greedily - the favoured warps eat
  // through the queue and the others hold their first trajectories to the end, so every chunk completes only when the
  // kernel does (OU/Heun: 4.45 ms in separate launches, 5.2 ms pipelined).  DFX_HOST_PIPE=2 forces the pipeline.
  const bool adaptive = h->controller == DFX_CTRL_PID;
  const char *e = std::getenv("DFX_HOST_PIPE");
  if (!(adaptive || (e && atoi(e) == 2))) return false;
  long long min_traj = 256 * 1024;
  if (const char *m = std::getenv("DFX_HOST_PIPE_MIN")) { const long long v = atoll(m); if (v >= 1024) min_traj = v; }  // experiments
  return !rich && h->field_id != DFX_FIELD_MLP && h->field_id < DFX_FIELD_USER && h->n_traj >= min_traj;
}

static int solve_host_pipelined(const dfx_solve_desc *h, int device) {
  const size_t es = h->dtype == DFX_F64 ? 8 : 4;
  const size_t N = (size_t)h->n_traj, D = (size_t)h->dim;
  const size_t T = (size_t)dfx_out_size(h);
  // chunk count: measured on C2 (B200, PCIe gen 5): 2^20 trajectories 16 chunks; 2^19: 12 (2.06 ms; 16: 2.10, 8: 2.08); 2^18: 8
  // (1.24 ms; 16: 1.30, 4: 1.26) - every chunk costs a completion round trip through mapped host memory
  int64_t want = h->n_traj >= (1 << 20) ? 16 : (h->n_traj >= (1 << 19) ? 12 : 8);
  if (const char *e = std::getenv("DFX_HOST_CHUNKS")) { const int64_t v = atoll(e); if (v >= 1) want = v; }
  if (want > kMaxPipeChunks) want = kMaxPipeChunks;
  const size_t chunk_len = ((N + (size_t)want - 1) / (size_t)want + 31) & ~(size_t)31;
  const int nchunks = (int)((N + chunk_len - 1) / chunk_len);

  // per-thread pinned control block: [0, 64) completion flags written by the kernel, [64, 128) the values 1..64 that
  // the copy engine moves into `in_ready`, [128] a non-zero word (copied into the device abort word when delivery fails)
  static thread_local unsigned *ctl_host = nullptr;
  if (!ctl_host) {
    DFX_CUDA_OK(cudaHostAlloc((void **)&ctl_host, (2 * kMaxPipeChunks + 1) * sizeof(unsigned), cudaHostAllocMapped | cudaHostAllocPortable));
    for (int i = 0; i < kMaxPipeChunks; ++i) ctl_host[kMaxPipeChunks + i] = (unsigned)(i + 1);
  }
  volatile unsigned *flags = ctl_host;
  for (int i = 0; i < nchunks; ++i) flags[i] = 0;
  ctl_host[2 * kMaxPipeChunks] = 0xFFFFFFFFu;
  unsigned *flags_dev = nullptr;
  DFX_CUDA_OK(cudaHostGetDevicePointer((void **)&flags_dev, ctl_host, 0));

  const bool trace = std::getenv("DFX_HOST_PIPE_TRACE") != nullptr;  // wall-clock marks of the pipeline on stderr
  const auto t_begin = std::chrono::steady_clock::now();
  auto now_ms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
  std::vector<double> t_flag;
  double t_enq = 0, t_launch = 0, t_k = 0, t_sync = 0;
  // streams and events are kept per host thread (creating them costs more than the first chunk's copy)
  struct PipeStreams { int device = -1; cudaStream_t s[3] = {nullptr, nullptr, nullptr}; cudaEvent_t ev[2] = {nullptr, nullptr}; };
  static thread_local PipeStreams ps;
  int rc = 0;
  auto fail = [&](const char *what, cudaError_t e) { if (!rc) { set_error("%s failed: %s", what, cudaGetErrorString(e)); rc = DFX_ERR_CUDA; } };
#define PIPE_OK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) fail(#expr, _e); } while (0)
  if (ps.device != device) {
    for (cudaStream_t &st : ps.s) { if (st) cudaStreamDestroy(st); st = nullptr; }
    for (cudaEvent_t &ev : ps.ev) { if (ev) cudaEventDestroy(ev); ev = nullptr; }
    for (cudaStream_t &st : ps.s) PIPE_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (cudaEvent_t &ev : ps.ev) PIPE_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    ps.device = rc ? -1 : device;
    if (rc) return rc;
  }
  cudaStream_t s_in = ps.s[0], s_k = ps.s[1], s_out = ps.s[2];
  cudaEvent_t ev_in = ps.ev[0], ev_k = ps.ev[1];
  // ONE stream-ordered allocation holds every device buffer (laid out twice: first to measure, then for real)
  char *block = nullptr;
  size_t cursor = 0;
  bool measuring = true;
  auto dev_alloc = [&](size_t bytes) -> char * {
    if (bytes == 0) return nullptr;
    const size_t at = cursor;
    cursor += (bytes + 255) & ~(size_t)255;
    return measuring ? (char *)(uintptr_t)256 : block + at;  // (non-null placeholder while measuring)
  };
  struct Slice { const char *host; char *dev; size_t per_traj; };
  std::vector<Slice> ins, outs;
  auto per_traj_in = [&](const void *src, size_t per) -> void * {
    if (!src) return nullptr;
    char *q = dev_alloc(N * per);
    ins.push_back({(const char *)src, q, per});
    return q;
  };
  auto per_traj_out = [&](void *dst, size_t per) -> void * {
    if (!dst || per == 0) return nullptr;
    char *q = dev_alloc(N * per);
    outs.push_back({(const char *)dst, q, per});
    return q;
  };
  auto whole_in = [&](const void *src, size_t bytes) -> void * {
    if (!src || bytes == 0) return nullptr;
    char *q = dev_alloc(bytes);
    if (!measuring) PIPE_OK(cudaMemcpyAsync(q, src, bytes, cudaMemcpyHostToDevice, s_in));
    return q;
  };
  dfx_solve_desc d = *h;
  unsigned *ctl_dev = nullptr;
  auto layout = [&] {
    cursor = 0;
    ins.clear();
    outs.clear();
    ctl_dev = (unsigned *)dev_alloc((2 + (size_t)nchunks) * sizeof(unsigned));  // [in_ready, abort, done[nchunks]]
    if (!measuring) PIPE_OK(cudaMemsetAsync(ctl_dev, 0, (2 + (size_t)nchunks) * sizeof(unsigned), s_in));
    d.field_weights = whole_in(h->field_weights, (size_t)h->n_field_weights * es);
    d.y0 = per_traj_in(h->y0, D * es);
    d.t0_per_traj = per_traj_in(h->t0_per_traj, es);
    d.t1_per_traj = per_traj_in(h->t1_per_traj, es);
    d.bm_keys = (const uint32_t *)per_traj_in(h->bm_keys, 8);
    d.traj_args = per_traj_in(h->traj_args, (size_t)h->n_traj_args * es);
    d.ts_out = per_traj_out(h->ts_out, T * es);
    d.ys_out = per_traj_out(h->ys_out, T * D * es);
    d.stats = (int32_t *)per_traj_out(h->stats, 12);
    d.result = (int32_t *)per_traj_out(h->result, 4);
    // totals: ONE launch, so the kernel accumulates straight into the caller's device words (or scratch for a host copy)
    d.totals = h->totals_device ? h->totals_device : (h->totals ? (int64_t *)dev_alloc(4 * sizeof(int64_t)) : nullptr);
    d.totals_device = nullptr;
    d.save_count = (int32_t *)per_traj_out(h->save_count, 4);
    // finals: the caller's device buffers when given (the host copies, if any, are taken from there chunk by chunk)
    if (h->y_final_device) { d.y_final = h->y_final_device; if (h->y_final) outs.push_back({(const char *)h->y_final, (char *)h->y_final_device, D * es}); }
    else d.y_final = per_traj_out(h->y_final, D * es);
    if (h->t_final_device) { d.t_final = h->t_final_device; if (h->t_final) outs.push_back({(const char *)h->t_final, (char *)h->t_final_device, es}); }
    else d.t_final = per_traj_out(h->t_final, es);
    d.y_final_device = d.t_final_device = nullptr;
  };
  layout();
  {
    const cudaError_t e = cudaMallocAsync((void **)&block, cursor, s_in);
    if (e != cudaSuccess) { fail("cudaMallocAsync", e); return rc; }
  }
  measuring = false;
  layout();
  // the kernel may start once the control words are zeroed and the replicated inputs are resident
  PIPE_OK(cudaEventRecord(ev_in, s_in));
  PIPE_OK(cudaStreamWaitEvent(s_k, ev_in, 0));
  auto enqueue_inputs = [&](int c) {
    const size_t lo = (size_t)c * chunk_len, cnt = std::min(chunk_len, N - lo);
    for (const Slice &sl : ins) PIPE_OK(cudaMemcpyAsync(sl.dev + lo * sl.per_traj, sl.host + lo * sl.per_traj, cnt * sl.per_traj, cudaMemcpyHostToDevice, s_in));
    PIPE_OK(cudaMemcpyAsync(ctl_dev, ctl_host + kMaxPipeChunks + c, sizeof(unsigned), cudaMemcpyHostToDevice, s_in));
  };
  // the first chunk, then the launch, then the other chunks: the kernel starts while they are still being enqueued
  if (!rc) enqueue_inputs(0);
  bool launched = false;
  t_enq = now_ms();
  if (!rc) {
    HostPipe hp{ctl_dev, ctl_dev + 2, flags_dev, (int)chunk_len, false};
    host_pipe() = &hp;
    rc = dfx_ensemble_solve(&d, (void *)s_k);
    host_pipe() = nullptr;
    if (!rc && !hp.consumed) { set_error("internal: the launcher ignored the host pipeline"); rc = DFX_ERR_UNSUPPORTED; }
    launched = rc == 0;
    if (launched) PIPE_OK(cudaEventRecord(ev_k, s_k));
  }
  t_launch = now_ms();
  if (const char *e = std::getenv("DFX_HOST_PIPE_FAULT")) {  // test hook: chunk delivery fails right after the launch
    if (launched && atoi(e) != 0) fail("injected fault", cudaErrorUnknown);
  }
  for (int c = 1; c < nchunks && launched && !rc; ++c) enqueue_inputs(c);
  // If anything failed after the launch, some chunks will never be delivered: release the kernel's waiting lanes by
  // setting the device abort word next to `in_ready` (on the output stream: the input stream may be the one that failed).
  bool abort_sent = false;
  auto release_kernel = [&] {
    if (!launched || abort_sent) return;
    abort_sent = true;
    cudaMemcpyAsync(ctl_dev + 1, ctl_host + 2 * kMaxPipeChunks, sizeof(unsigned), cudaMemcpyHostToDevice, s_out);
  };
  if (rc) release_kernel();
  const double t_enq_all = now_ms();
  std::vector<unsigned> v_flag;
  // chunks complete roughly, not exactly, in order: every pass sends whichever have become ready
  std::vector<char> sent((size_t)nchunks, 0);
  int remaining = launched && !rc ? nchunks : 0;
  bool kernel_done = false;
  for (unsigned spin = 1; remaining > 0 && !rc; ++spin) {
    bool progress = false;
    for (int c = 0; c < nchunks; ++c) {
      if (sent[c] || flags[c] == 0) continue;
      if (trace) { t_flag.push_back(now_ms()); v_flag.push_back((unsigned)flags[c]); }
      const size_t lo = (size_t)c * chunk_len, cnt = std::min(chunk_len, N - lo);
      for (const Slice &sl : outs) PIPE_OK(cudaMemcpyAsync((char *)sl.host + lo * sl.per_traj, sl.dev + lo * sl.per_traj, cnt * sl.per_traj, cudaMemcpyDeviceToHost, s_out));
      sent[c] = 1;
      --remaining;
      progress = true;
    }
    if (progress || remaining == 0) continue;
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();  // polite busy-wait: the flags arrive over PCIe, microseconds apart at best
#endif
    if (kernel_done) { set_error("internal: %d chunks were never completed", remaining); rc = DFX_ERR_CUDA; break; }
    if ((spin & 1023) == 0) {  // a finished or failed kernel ends the wait (one more pass picks up its last flags)
      const cudaError_t q = cudaEventQuery(ev_k);
      if (q == cudaSuccess) kernel_done = true;
      else if (q != cudaErrorNotReady) fail("ensemble kernel", q);
      const cudaError_t qi = cudaStreamQuery(s_in);  // an input copy that failed asynchronously never bumps in_ready
      if (qi != cudaSuccess && qi != cudaErrorNotReady) fail("input copies", qi);
    }
  }
  if (rc) release_kernel();  // never leave the kernel waiting for inputs
  if (launched && !rc && h->totals && d.totals)  // stream-ordered after the kernel
    PIPE_OK(cudaMemcpyAsync(h->totals, d.totals, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost, s_k));
  if (trace && s_k) { cudaStreamSynchronize(s_k); t_k = now_ms(); }
  for (cudaStream_t st : {s_in, s_k, s_out})
    if (st) { const cudaError_t e = cudaStreamSynchronize(st); if (e != cudaSuccess) fail("stream sync", e); }
  t_sync = now_ms();
  if (trace) {
    std::fprintf(stderr, "[dfx host pipe] %d chunks of %zu: first H2D enqueued %.3f ms, launched %.3f, all H2D enqueued %.3f, chunk flags seen at", nchunks, chunk_len, t_enq, t_launch, t_enq_all);
    for (double t : t_flag) std::fprintf(stderr, " %.3f", t);
    std::fprintf(stderr, " (device clock, ms after the first:");
    for (unsigned v : v_flag) std::fprintf(stderr, " %.3f", (double)(int)(v - v_flag[0]) * 1.024e-3);
    std::fprintf(stderr, ")");
    std::fprintf(stderr, ", kernel done %.3f, all copies done %.3f\n", t_k, t_sync);
  }
  cudaFreeAsync(block, s_in);  // (every stream that touched the block has been synchronised)
#undef PIPE_OK
  return rc;
}

int dfx_ensemble_solve_host(const dfx_solve_desc *h, int device) {
  if (int rc = check_desc(h)) return rc;
  if (h->n_traj == 0) { if (h->totals) std::memset(h->totals, 0, 4 * sizeof(int64_t)); return 0; }
  if (dfx_device_count() <= device) { set_error("CUDA device %d not available", device); return DFX_ERR_NO_DEVICE; }
  DFX_CUDA_OK(cudaSetDevice(device));
  if (host_pipe_eligible(h)) return solve_host_pipelined(h, device);
  // chunks of >= 128K trajectories keep every SM busy; at most 8 chunks; dense output stays in one piece
  int64_t nchunks = h->save_dense ? 1 : h->n_traj / (128 * 1024);
  if (nchunks < 1) nchunks = 1;
  if (nchunks > 8) nchunks = 8;
  if (const char *e = std::getenv("DFX_HOST_CHUNKS")) {  // experiments
    const int64_t v = atoll(e);
    if (v >= 1 && !h->save_dense) nchunks = v > h->n_traj ? (h->n_traj > 0 ? h->n_traj : 1) : v;
  }
  // two non-blocking streams per host thread and device, created once (creating / destroying them costs more than the
  // copies of a small ensemble)
  struct ChunkStreams { int device = -1; cudaStream_t s[2] = {nullptr, nullptr}; };
  static thread_local ChunkStreams cs;
  if (cs.device != device) {
    for (cudaStream_t &x : cs.s) { if (x) cudaStreamDestroy(x); x = nullptr; }
    for (cudaStream_t &x : cs.s) DFX_CUDA_OK(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
    cs.device = device;
  }
  cudaStream_t st[2] = {cs.s[0], cs.s[1]};
  const int nstreams = nchunks > 1 ? 2 : 1;
  std::vector<void *> allocs[2];
  const bool want_totals = h->totals || h->totals_device;
  // per-chunk totals land in PINNED memory: a D2H copy into pageable memory would block the host and serialise the chunks
  static thread_local int64_t *chunk_totals = nullptr;
  if (want_totals) {
    if (!chunk_totals) DFX_CUDA_OK(cudaHostAlloc((void **)&chunk_totals, 4 * 64 * sizeof(int64_t), cudaHostAllocPortable));
    if (nchunks > 64) nchunks = 64;
    std::memset(chunk_totals, 0, 4 * (size_t)nchunks * sizeof(int64_t));
  }
  int rc = 0;
  for (int64_t c = 0; c < nchunks && !rc; ++c) {
    const int64_t lo = h->n_traj * c / nchunks, hi = h->n_traj * (c + 1) / nchunks;
    rc = solve_host_chunk(h, lo, hi - lo, st[c % nstreams], allocs[c % nstreams], want_totals ? &chunk_totals[4 * (size_t)c] : nullptr);
  }
  for (int i = 0; i < nstreams; ++i) {
    for (void *p : allocs[i]) cudaFreeAsync(p, st[i]);
    cudaError_t e = cudaStreamSynchronize(st[i]);
    if (!rc && e != cudaSuccess) { set_error("stream sync failed: %s", cudaGetErrorString(e)); rc = DFX_ERR_CUDA; }
  }
  if (want_totals && !rc && !(nchunks == 1 && h->totals_device)) {  // combine the chunks' totals: sums, and the max of the per-trajectory maxima
    int64_t tot[4] = {0, 0, 0, 0};
    for (int64_t c = 0; c < nchunks; ++c) {
      for (int k = 0; k < 3; ++k) tot[k] += chunk_totals[4 * (size_t)c + k];
      tot[3] = std::max(tot[3], chunk_totals[4 * (size_t)c + 3]);
    }
    if (h->totals) std::memcpy(h->totals, tot, sizeof(tot));
    if (h->totals_device) {
      const cudaError_t e = cudaMemcpy(h->totals_device, tot, sizeof(tot), cudaMemcpyHostToDevice);
      if (e != cudaSuccess) { set_error("totals H2D failed: %s", cudaGetErrorString(e)); rc = DFX_ERR_CUDA; }
    }
  }
  return rc;
}

int dfx_peer_alloc(int64_t bytes, void **device_ptr, void *ipc_handle) {
  if (bytes <= 0 || !device_ptr || !ipc_handle) { set_error("peer_alloc: bad argument"); return DFX_ERR_BAD_ARGUMENT; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  DFX_CUDA_OK(cudaMalloc(device_ptr, (size_t)bytes));
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, *device_ptr);
  if (e != cudaSuccess) { cudaFree(*device_ptr); *device_ptr = nullptr; set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); return DFX_ERR_CUDA; }
  std::memcpy(ipc_handle, &h, 64);
  return 0;
}
int dfx_peer_open(const void *ipc_handle, void **device_ptr) {
  if (!ipc_handle || !device_ptr) { set_error("peer_open: bad argument"); return DFX_ERR_BAD_ARGUMENT; }
  cudaIpcMemHandle_t h;
  std::memcpy(&h, ipc_handle, 64);
  DFX_CUDA_OK(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
int dfx_peer_close(void *device_ptr) { if (device_ptr) DFX_CUDA_OK(cudaIpcCloseMemHandle(device_ptr)); return 0; }
int dfx_peer_free(void *device_ptr) { if (device_ptr) DFX_CUDA_OK(cudaFree(device_ptr)); return 0; }

int dfx_broadcast_device_scalar(int dtype, int64_t n, const void *src_device, void *dst_device, void *stream) {
  if (n <= 0) return 0;
  if (!src_device || !dst_device) { set_error("broadcast: null pointer"); return DFX_ERR_BAD_ARGUMENT; }
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (dtype == DFX_F64) broadcast_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(n, (const double *)src_device, (double *)dst_device);
  else broadcast_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(n, (const float *)src_device, (float *)dst_device);
  count_launch();
  DFX_CUDA_OK(cudaGetLastError());
  return 0;
}

int dfx_threefry2x32(int64_t n, const uint32_t *keys, const uint32_t *ctrs, uint32_t *out, void *stream) {
  if (n <= 0) return 0;
  threefry_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, keys, ctrs, out);
  count_launch();
  DFX_CUDA_OK(cudaGetLastError());
  return 0;
}
int dfx_random_split(int64_t n, const uint32_t *keys, int num, int partitionable, uint32_t *out, void *stream) {
  if (n <= 0) return 0;
  split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, keys, num, partitionable, out);
  count_launch();
  DFX_CUDA_OK(cudaGetLastError());
  return 0;
}
int dfx_random_normal(int dtype, int64_t n, const uint32_t *keys, int partitionable, void *out, int m, void *stream) {
  if (n <= 0) return 0;
  if (m < 0 || m > kMaxDim) { set_error("normal: shape (m,) with 0 <= m <= %d (0 = scalar), got %d", kMaxDim, m); return DFX_ERR_BAD_ARGUMENT; }
  cudaStream_t st = (cudaStream_t)stream;
#define DFX_NORMAL_CASE(M) case M: if (dtype == DFX_F64) launch_normal<double, M>(n, keys, partitionable, out, st); else launch_normal<float, M>(n, keys, partitionable, out, st); break;
  switch (m == 0 ? 1 : m) {
    DFX_NORMAL_CASE(1) DFX_NORMAL_CASE(2) DFX_NORMAL_CASE(3) DFX_NORMAL_CASE(4)
    DFX_NORMAL_CASE(5) DFX_NORMAL_CASE(6) DFX_NORMAL_CASE(7) DFX_NORMAL_CASE(8)
  }
#undef DFX_NORMAL_CASE
  count_launch();
  DFX_CUDA_OK(cudaGetLastError());
  return 0;
}

int dfx_vbt_evaluate(int dtype, int levy_area, int partitionable, int64_t n, const uint32_t *keys, double bm_t0,
                     double bm_t1, double bm_tol, const void *ta, const void *tb, int per_traj_times, void *W, void *H,
                     int bm_dim, void *stream) {
  if (n <= 0) return 0;
  if (!(bm_t0 < bm_t1)) { set_error("t0 must be strictly less than t1"); return DFX_ERR_BAD_ARGUMENT; }
  if (bm_dim < 0 || bm_dim > kMaxDim) { set_error("VirtualBrownianTree shape () or (m,) with m <= %d, got m = %d", kMaxDim, bm_dim); return DFX_ERR_BAD_ARGUMENT; }
  VbtParams vp;
  vp.t0 = bm_t0; vp.t1 = bm_t1; vp.levy = levy_area; vp.partitionable = partitionable;
  vp.cache_levels = 0; vp.cache_stride = 0;
  const double tol_n = bm_tol / (bm_t1 - bm_t0);
  int depth = 0;
  while (std::ldexp(1.0, -depth) > tol_n && depth < 1000) ++depth;
  vp.depth = depth;
  cudaStream_t st = (cudaStream_t)stream;
  const bool stla = levy_area == DFX_LEVY_SPACE_TIME;
#define DFX_VBT_CASE(M) case M: if (dtype == DFX_F64) launch_vbt<double, M>(stla, n, keys, vp, ta, tb, per_traj_times, W, H, st); else launch_vbt<float, M>(stla, n, keys, vp, ta, tb, per_traj_times, W, H, st); break;
  switch (bm_dim == 0 ? 1 : bm_dim) {
    DFX_VBT_CASE(1) DFX_VBT_CASE(2) DFX_VBT_CASE(3) DFX_VBT_CASE(4)
    DFX_VBT_CASE(5) DFX_VBT_CASE(6) DFX_VBT_CASE(7) DFX_VBT_CASE(8)
  }
#undef DFX_VBT_CASE
  count_launch();
  DFX_CUDA_OK(cudaGetLastError());
  return 0;
}

static int dense_eval_entry(bool deriv, int dtype, int solver_id, int64_t n_traj, int dim, int max_steps, const void *dense_ts,
                            const void *dense_y0, const void *dense_y1, const void *dense_k, const int32_t *dense_count,
                            double direction, const void *tq, int nq, void *out, void *stream) {
  if (n_traj <= 0 || nq <= 0) return 0;
  if (dtype == DFX_F64)
    return dense_eval_dispatch<double>(deriv, solver_id, dim, n_traj, max_steps, dense_ts, dense_y0, dense_y1, dense_k,
                                       dense_count, direction, tq, nq, out, (cudaStream_t)stream);
  return dense_eval_dispatch<float>(deriv, solver_id, dim, n_traj, max_steps, dense_ts, dense_y0, dense_y1, dense_k,
                                    dense_count, direction, tq, nq, out, (cudaStream_t)stream);
}
int dfx_dense_evaluate(int dtype, int solver_id, int64_t n_traj, int dim, int max_steps, const void *dense_ts,
                       const void *dense_y0, const void *dense_y1, const void *dense_k, const int32_t *dense_count,
                       double direction, const void *tq, int nq, void *out, void *stream) {
  return dense_eval_entry(false, dtype, solver_id, n_traj, dim, max_steps, dense_ts, dense_y0, dense_y1, dense_k, dense_count,
                          direction, tq, nq, out, stream);
}
int dfx_dense_derivative(int dtype, int solver_id, int64_t n_traj, int dim, int max_steps, const void *dense_ts,
                         const void *dense_y0, const void *dense_y1, const void *dense_k, const int32_t *dense_count,
                         double direction, const void *tq, int nq, void *out, void *stream) {
  return dense_eval_entry(true, dtype, solver_id, n_traj, dim, max_steps, dense_ts, dense_y0, dense_y1, dense_k, dense_count,
                          direction, tq, nq, out, stream);
}

int dfx_dense_pad(int dtype, int solver_id, int64_t n_traj, int dim, int max_steps, void *dense_ts, void *dense_y0, void *dense_y1,
                  void *dense_k, const int32_t *dense_count, void *stream) {
  if (n_traj <= 0 || max_steps <= 0) return 0;
  if (!dense_ts || !dense_y0 || !dense_y1 || !dense_count) { set_error("dense_pad: null buffer"); return DFX_ERR_BAD_ARGUMENT; }
  const int s = dfx_num_stages(solver_id);
  if (s < 0) { set_error("unknown solver %d", solver_id); return DFX_ERR_BAD_ARGUMENT; }
  const unsigned blocks = (unsigned)((n_traj * 32 + 255) / 256);
  if (dtype == DFX_F64) dense_pad_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(n_traj, max_steps, dim, s * dim, (double *)dense_ts, (double *)dense_y0, (double *)dense_y1, (double *)dense_k, dense_count);
  else dense_pad_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(n_traj, max_steps, dim, s * dim, (float *)dense_ts, (float *)dense_y0, (float *)dense_y1, (float *)dense_k, dense_count);
  count_launch();
  DFX_CUDA_OK(cudaGetLastError());
  return 0;
}

double dfx_measure_fma_peak(int dtype, int device) {
  if (dfx_device_count() <= device) { set_error("CUDA device %d not available", device); return DFX_ERR_NO_DEVICE; }
  cudaSetDevice(device);
  int sms = 0;
  if (device_sm_count(&sms)) return DFX_ERR_CUDA;
  const int threads = 256, blocks = sms * 8, iters = dtype == DFX_F64 ? 4096 : 16384;
  void *out = nullptr;
  if (cudaMalloc(&out, 64) != cudaSuccess) return DFX_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    if (dtype == DFX_F64) fma_peak_kernel<double><<<blocks, threads>>>((double *)out, iters, 1.0000001, 1e-9);
    else fma_peak_kernel<float><<<blocks, threads>>>((float *)out, iters, 1.0000001f, 1e-9f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  count_launch(6);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (cudaGetLastError() != cudaSuccess) return DFX_ERR_CUDA;
  const double flops = 2.0 * 8.0 * (double)iters * (double)threads * (double)blocks;
  return flops / (best * 1e-3) / 1e12;
}

double dfx_measure_int_peak(int device) {
  if (dfx_device_count() <= device) { set_error("CUDA device %d not available", device); return DFX_ERR_NO_DEVICE; }
  cudaSetDevice(device);
  int sms = 0;
  if (device_sm_count(&sms)) return DFX_ERR_CUDA;
  const int threads = 256, blocks = sms * 8, iters = 8192;
  void *out = nullptr;
  if (cudaMalloc(&out, 64) != cudaSuccess) return DFX_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    int_peak_kernel<<<blocks, threads>>>((uint32_t *)out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  count_launch(6);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (cudaGetLastError() != cudaSuccess) return DFX_ERR_CUDA;
  const double ops = 3.0 * 8.0 * (double)iters * (double)threads * (double)blocks;  // add, rotate, xor
  return ops / (best * 1e-3) / 1e12;
}

}  // extern "C"
