// xla_ffi_shim.cc - XLA FFI (jax.ffi) handlers over the plain C ABI of include/diffrax_b200.h.
//
// The boundary the north star names: "the Python host dispatches through a thin jax.ffi C-ABI custom call".  The
// handlers below re-pack XLA's buffers + attributes into dfx_solve_desc and call dfx_ensemble_solve on XLA's stream; all
// numerics stay in libdiffrax_b200.so.  What they replace on the reference side is everything `diffeqsolve`
// (/root/reference/diffrax/_integrate.py:890-911) does from `adjoint.loop` (:1456) down.
//
// Batching.  The custom call is written for ONE trajectory - y0 [d], t0 [], t1 [], key [2] -> ts [T], ys [T, d],
// stats [3], result [] ... - and registered with vmap_method="expand_dims" (INTEGRATION.md): under
// `jax.vmap(diffeqsolve)` XLA hands the handler the same operands with a leading batch axis (N for batched operands,
// 1 for unbatched ones), and the handler launches the ensemble kernel ONCE for the whole batch.  Operand conventions:
//   y0        [d] or [N, d]
//   t0s, t1s  [0] (use the static `t0` / `t1` attributes), [N] (per-trajectory regions), or one element on the device
//             (a traced scalar that vmap did not batch: broadcast on the stream)
//   save_ts   [T] or [1, T] (shared by all trajectories; T may be 0), step_ts / jump_ts likewise
//   keys      [2] or [N, 2] uint32 key data (zero elements for ODEs)
//   state_in  [0] or [N, 5 + d];  field_weights [0] or the MLP weights
//   traj_args [0], or [K] / [N, K]: the vmapped `args` of diffeqsolve (per-trajectory functor parameters; an unbatched [K]
//             under a batch of N > 1 is the same as passing it through the `params` attribute and is rejected here)
// Results are caller(XLA)-owned; unused ones (dense_* without SaveAt(dense), state_out without SaveAt(solver_state=...))
// are declared with zero elements by the Python side.
//
// Build: `python -m diffrax_b200.build --ffi` compiles this file against jaxlib's header when `jax.ffi.include_dir()`
// exists, else against the compile-check stub tests/ffi_stub (tests/test_ffi_shim.py does that on every run).
#include <cmath>
#include <cstring>
#include <string>

#include <cuda_runtime_api.h>

#include "diffrax_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

template <ffi::DataType DT>
ffi::Error EnsembleSolve(cudaStream_t stream,
                         // ---- operands ----
                         ffi::Buffer<DT> y0, ffi::Buffer<DT> t0s, ffi::Buffer<DT> t1s, ffi::Buffer<DT> save_ts,
                         ffi::Buffer<DT> step_ts, ffi::Buffer<DT> jump_ts, ffi::Buffer<ffi::U32> keys,
                         ffi::Buffer<DT> state_in, ffi::Buffer<DT> field_weights, ffi::Buffer<DT> traj_args,
                         // ---- results ----
                         ffi::ResultBuffer<DT> ts_out, ffi::ResultBuffer<DT> ys_out, ffi::ResultBuffer<ffi::S32> stats,
                         ffi::ResultBuffer<ffi::S32> result, ffi::ResultBuffer<DT> y_final, ffi::ResultBuffer<DT> t_final,
                         ffi::ResultBuffer<DT> dense_ts, ffi::ResultBuffer<DT> dense_y0, ffi::ResultBuffer<DT> dense_y1,
                         ffi::ResultBuffer<DT> dense_k, ffi::ResultBuffer<ffi::S32> dense_count,
                         ffi::ResultBuffer<DT> state_out,
                         // ---- attributes ----
                         int32_t field_id, int32_t solver_id, int32_t controller, int32_t levy_area, int32_t bm_dim,
                         double t0, double t1, double dt0, double rtol, double atol, double pcoeff, double icoeff,
                         double dcoeff, double safety, double factormin, double factormax, double dtmin, double dtmax,
                         int32_t force_dtmin, double error_order, int32_t hairer_initial_step,
                         int32_t store_rejected_steps, int32_t save_t0, int32_t save_t1, int32_t save_steps,
                         int32_t save_dense, int32_t max_steps, double bm_t0, double bm_t1, double bm_tol,
                         int32_t partitionable, int32_t state_in_flags, ffi::Span<const int32_t> event_kind,
                         ffi::Span<const int32_t> event_direction, int32_t event_root_find, double event_rtol,
                         double event_atol, ffi::Span<const double> event_params, ffi::Span<const double> params) {
  dfx_solve_desc d;
  std::memset(&d, 0, sizeof d);
  d.struct_size = sizeof d;
  d.abi_version = DFX_ABI_VERSION;
  const auto dims = y0.dimensions();
  if (dims.size() < 1 || dims.size() > 2) return ffi::Error::InvalidArgument("y0 must be [d] or [N, d]");
  const int64_t n = dims.size() == 2 ? dims[0] : 1;
  d.n_traj = n;
  d.dim = static_cast<int32_t>(dims.back());
  d.dtype = DT == ffi::F64 ? DFX_F64 : DFX_F32;
  const size_t es = DT == ffi::F64 ? 8 : 4;
  d.field_id = field_id; d.solver_id = solver_id; d.controller = controller; d.levy_area = levy_area; d.bm_dim = bm_dim;
  d.field_params = params.begin(); d.n_field_params = static_cast<int32_t>(params.size());
  if (field_weights.element_count()) { d.field_weights = field_weights.untyped_data(); d.n_field_weights = (int64_t)field_weights.element_count(); }
  d.y0 = y0.untyped_data();
  d.t0 = t0; d.t1 = t1; d.dt0 = dt0;
  if (traj_args.element_count()) {
    if (traj_args.element_count() % (size_t)n != 0) return ffi::Error::InvalidArgument("traj_args must be [K] or [N, K] with N = the batch of y0");
    d.traj_args = traj_args.untyped_data();
    d.n_traj_args = static_cast<int32_t>(traj_args.element_count() / (size_t)n);
  }

  // per-trajectory integration regions ("vmappable everything, including the region of integration", README.md:10)
  void *scratch = nullptr;
  auto region = [&](const ffi::Buffer<DT> &b, int slot, const void **out) -> const char * {
    const int64_t m = static_cast<int64_t>(b.element_count());
    if (m == 0) { *out = nullptr; return nullptr; }                       // static attribute
    if (m == n) { *out = b.untyped_data(); return nullptr; }              // one value per trajectory
    if (m != 1) return "t0 / t1 must have 0, 1 or N elements";
    if (!scratch && cudaMallocAsync(&scratch, 2 * (size_t)n * es, stream) != cudaSuccess) return "cudaMallocAsync failed";
    void *dst = static_cast<char *>(scratch) + (size_t)slot * (size_t)n * es;  // a traced scalar: broadcast on the stream
    if (dfx_broadcast_device_scalar(d.dtype, n, b.untyped_data(), dst, static_cast<void *>(stream)) != DFX_OK) return dfx_last_error();
    *out = dst;
    return nullptr;
  };
  if (const char *e = region(t0s, 0, &d.t0_per_traj)) return ffi::Error::InvalidArgument(e);
  if (const char *e = region(t1s, 1, &d.t1_per_traj)) { if (scratch) cudaFreeAsync(scratch, stream); return ffi::Error::InvalidArgument(e); }
  if ((d.t0_per_traj == nullptr) != (d.t1_per_traj == nullptr)) {
    if (scratch) cudaFreeAsync(scratch, stream);
    return ffi::Error::InvalidArgument("pass t0 and t1 either both as attributes or both as operands");
  }

  d.rtol = rtol; d.atol = atol; d.pcoeff = pcoeff; d.icoeff = icoeff; d.dcoeff = dcoeff; d.safety = safety;
  d.factormin = factormin; d.factormax = factormax; d.dtmin = dtmin; d.dtmax = dtmax; d.force_dtmin = force_dtmin;
  d.error_order = error_order; d.hairer_initial_step = hairer_initial_step;
  auto shared_times = [&](const ffi::Buffer<DT> &b, const void **ptr, int32_t *count) -> bool {
    const auto bd = b.dimensions();
    const int64_t last = bd.size() ? bd.back() : 0, total = static_cast<int64_t>(b.element_count());
    if (total != last) return false;                                      // a batched (per-trajectory) time grid
    *count = static_cast<int32_t>(last);
    *ptr = last ? b.untyped_data() : nullptr;
    return true;
  };
  bool ok = shared_times(save_ts, &d.save_ts, &d.n_save_ts) && shared_times(step_ts, &d.step_ts, &d.n_step_ts) &&
            shared_times(jump_ts, &d.jump_ts, &d.n_jump_ts);
  if (!ok) {
    if (scratch) cudaFreeAsync(scratch, stream);
    return ffi::Error::InvalidArgument("saveat.ts / step_ts / jump_ts are shared by the batch: do not vmap over them");
  }
  d.store_rejected_steps = store_rejected_steps;
  d.save_t0 = save_t0; d.save_t1 = save_t1; d.save_steps = save_steps; d.save_dense = save_dense; d.max_steps = max_steps;

  d.ts_out = ts_out->element_count() ? ts_out->untyped_data() : nullptr;
  d.ys_out = ys_out->element_count() ? ys_out->untyped_data() : nullptr;
  d.stats = stats->typed_data(); d.result = result->typed_data();
  d.y_final = y_final->element_count() ? y_final->untyped_data() : nullptr;
  d.t_final = t_final->element_count() ? t_final->untyped_data() : nullptr;
  if (save_dense) {
    d.dense_ts = dense_ts->untyped_data(); d.dense_y0 = dense_y0->untyped_data(); d.dense_y1 = dense_y1->untyped_data();
    d.dense_k = dense_k->element_count() ? dense_k->untyped_data() : nullptr;
    d.dense_count = dense_count->typed_data();
  }
  if (state_in.element_count()) { d.state_in = state_in.untyped_data(); d.state_in_flags = state_in_flags; }
  if (state_out->element_count()) d.state_out = state_out->untyped_data();

  if (levy_area != DFX_LEVY_NONE) {
    if (static_cast<int64_t>(keys.element_count()) != 2 * n) {
      if (scratch) cudaFreeAsync(scratch, stream);
      return ffi::Error::InvalidArgument("VirtualBrownianTree keys must be [N, 2] key data: vmap over the key");
    }
    d.bm_keys = keys.typed_data();
  }
  d.bm_t0 = bm_t0; d.bm_t1 = bm_t1; d.bm_tol = bm_tol; d.threefry_partitionable = partitionable;

  d.n_events = static_cast<int32_t>(event_kind.size());
  if (d.n_events > DFX_MAX_EVENTS || event_direction.size() != event_kind.size()) {
    if (scratch) cudaFreeAsync(scratch, stream);
    return ffi::Error::InvalidArgument("at most 4 event conditions, one direction each");
  }
  for (int i = 0; i < d.n_events; ++i) { d.event_kind[i] = event_kind[i]; d.event_direction[i] = event_direction[i]; }
  d.event_root_find = event_root_find; d.event_rtol = event_rtol; d.event_atol = event_atol;
  d.event_params = event_params.size() ? event_params.begin() : nullptr;
  d.n_event_params = static_cast<int32_t>(event_params.size());

  const int rc = dfx_ensemble_solve(&d, static_cast<void *>(stream));
  if (scratch) cudaFreeAsync(scratch, stream);
  if (rc != DFX_OK) return ffi::Error(rc == DFX_ERR_BAD_ARGUMENT ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                                      std::string(dfx_last_error()));
  return ffi::Error::Success();
}

// DenseInterpolation.evaluate / .derivative (_global_interpolation.py:335-368) on the buffers a SaveAt(dense=True) call returned
template <ffi::DataType DT>
ffi::Error DenseEvaluate(cudaStream_t stream, ffi::Buffer<DT> dense_ts, ffi::Buffer<DT> dense_y0, ffi::Buffer<DT> dense_y1,
                         ffi::Buffer<DT> dense_k, ffi::Buffer<ffi::S32> dense_count, ffi::Buffer<DT> tq,
                         ffi::ResultBuffer<DT> out, int32_t solver_id, int32_t derivative, double direction) {
  const auto td = dense_ts.dimensions(), yd = dense_y0.dimensions(), qd = tq.dimensions();
  if (td.size() != 2 || yd.size() != 3 || qd.size() != 2 || qd[0] != td[0]) return ffi::Error::InvalidArgument("dense_ts [N, max_steps + 1], dense_y0 [N, max_steps, d], tq [N, nq]");
  const int64_t n = td[0];
  const int max_steps = static_cast<int>(td[1] - 1), dim = static_cast<int>(yd[2]), nq = static_cast<int>(qd[1]);
  const void *k = dense_k.element_count() ? dense_k.untyped_data() : nullptr;
  const int dtype = DT == ffi::F64 ? DFX_F64 : DFX_F32;
  const int rc = (derivative ? dfx_dense_derivative : dfx_dense_evaluate)(
      dtype, solver_id, n, dim, max_steps, dense_ts.untyped_data(), dense_y0.untyped_data(), dense_y1.untyped_data(), k,
      dense_count.typed_data(), direction, tq.untyped_data(), nq, out->untyped_data(), static_cast<void *>(stream));
  if (rc != DFX_OK) return ffi::Error::InvalidArgument(std::string(dfx_last_error()));
  return ffi::Error::Success();
}

}  // namespace

#define DFX_BIND_SOLVE(DT)                                                                                          \
  ffi::Ffi::Bind()                                                                                                  \
      .Ctx<ffi::PlatformStream<cudaStream_t>>()                                                                     \
      .Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<DT>>()                   \
      .Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<ffi::U32>>().Arg<ffi::Buffer<DT>>()             \
      .Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<DT>>()                                                                \
      .Ret<ffi::Buffer<DT>>().Ret<ffi::Buffer<DT>>().Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::S32>>()       \
      .Ret<ffi::Buffer<DT>>().Ret<ffi::Buffer<DT>>()                                                                \
      .Ret<ffi::Buffer<DT>>().Ret<ffi::Buffer<DT>>().Ret<ffi::Buffer<DT>>().Ret<ffi::Buffer<DT>>()                   \
      .Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<DT>>()                                                          \
      .Attr<int32_t>("field_id").Attr<int32_t>("solver_id").Attr<int32_t>("controller").Attr<int32_t>("levy_area")  \
      .Attr<int32_t>("bm_dim")                                                                                      \
      .Attr<double>("t0").Attr<double>("t1").Attr<double>("dt0").Attr<double>("rtol").Attr<double>("atol")           \
      .Attr<double>("pcoeff").Attr<double>("icoeff").Attr<double>("dcoeff").Attr<double>("safety")                  \
      .Attr<double>("factormin").Attr<double>("factormax").Attr<double>("dtmin").Attr<double>("dtmax")               \
      .Attr<int32_t>("force_dtmin").Attr<double>("error_order").Attr<int32_t>("hairer_initial_step")                \
      .Attr<int32_t>("store_rejected_steps").Attr<int32_t>("save_t0").Attr<int32_t>("save_t1")                      \
      .Attr<int32_t>("save_steps").Attr<int32_t>("save_dense").Attr<int32_t>("max_steps")                           \
      .Attr<double>("bm_t0").Attr<double>("bm_t1").Attr<double>("bm_tol").Attr<int32_t>("partitionable")             \
      .Attr<int32_t>("state_in_flags").Attr<ffi::Span<const int32_t>>("event_kind")                                 \
      .Attr<ffi::Span<const int32_t>>("event_direction").Attr<int32_t>("event_root_find")                           \
      .Attr<double>("event_rtol").Attr<double>("event_atol").Attr<ffi::Span<const double>>("event_params")          \
      .Attr<ffi::Span<const double>>("params")

#define DFX_BIND_DENSE(DT)                                                                                          \
  ffi::Ffi::Bind()                                                                                                  \
      .Ctx<ffi::PlatformStream<cudaStream_t>>()                                                                     \
      .Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<DT>>()                   \
      .Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<DT>>()                                                          \
      .Ret<ffi::Buffer<DT>>()                                                                                       \
      .Attr<int32_t>("solver_id").Attr<int32_t>("derivative").Attr<double>("direction")

XLA_FFI_DEFINE_HANDLER_SYMBOL(DfxEnsembleSolveF64, EnsembleSolve<ffi::F64>, DFX_BIND_SOLVE(ffi::F64));
XLA_FFI_DEFINE_HANDLER_SYMBOL(DfxEnsembleSolveF32, EnsembleSolve<ffi::F32>, DFX_BIND_SOLVE(ffi::F32));
XLA_FFI_DEFINE_HANDLER_SYMBOL(DfxDenseEvaluateF64, DenseEvaluate<ffi::F64>, DFX_BIND_DENSE(ffi::F64));
XLA_FFI_DEFINE_HANDLER_SYMBOL(DfxDenseEvaluateF32, DenseEvaluate<ffi::F32>, DFX_BIND_DENSE(ffi::F32));
