// xla_ffi_shim.cc - thin XLA FFI (jax.ffi) handler over the plain C ABI of include/diffrax_b200.h.
//
// NOT part of the default build: the XLA FFI headers ship with jaxlib (jax.ffi.include_dir()) and
// jax is not installable in the authoring container.  Where jax is available:
//
//   g++ -O2 -std=c++17 -fPIC -shared -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -I include -I /usr/local/cuda/include diffrax_b200/csrc/ffi/xla_ffi_shim.cc \
//       -L diffrax_b200/lib -ldiffrax_b200 -Wl,-rpath,'$ORIGIN' -o diffrax_b200/lib/libdfx_xla_ffi.so
//
// The handler only re-packs buffers + attributes into dfx_solve_desc and calls dfx_ensemble_solve
// on XLA's stream; all numerics stay in libdiffrax_b200.so.  See INTEGRATION.md for the Python side.
#include <cmath>
#include <cstring>
#include <string>

#include <cuda_runtime_api.h>

#include "diffrax_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

template <ffi::DataType DT>
static ffi::Error EnsembleSolve(cudaStream_t stream,
                                ffi::Buffer<DT> y0,                 // [N, d]
                                ffi::Buffer<DT> t0t1,               // [2] or [N, 2]: per-trajectory region when rank 2
                                ffi::Buffer<DT> save_ts,            // [T] (T may be 0)
                                ffi::Buffer<ffi::U32> keys,         // [N, 2] (N may be 0 for ODEs)
                                ffi::Buffer<ffi::F64> field_params, // HOST-visible? no: attributes below carry them
                                ffi::ResultBuffer<DT> ts_out,       // [N, T_out]
                                ffi::ResultBuffer<DT> ys_out,       // [N, T_out, d]
                                ffi::ResultBuffer<ffi::S32> stats,  // [N, 3]
                                ffi::ResultBuffer<ffi::S32> result, // [N]
                                ffi::ResultBuffer<DT> y_final,      // [N, d]
                                int32_t field_id, int32_t solver_id, int32_t controller, int32_t levy_area,
                                double t0, double t1, double dt0, double rtol, double atol, double pcoeff,
                                double icoeff, double dcoeff, double safety, double factormin, double factormax,
                                double dtmin, double dtmax, int32_t force_dtmin, double error_order,
                                int32_t save_t0, int32_t save_t1, int32_t save_steps, int32_t max_steps,
                                double bm_t0, double bm_t1, double bm_tol, int32_t partitionable,
                                ffi::Span<const double> params) {
  (void)field_params;
  (void)t0t1;
  dfx_solve_desc d;
  std::memset(&d, 0, sizeof d);
  d.struct_size = sizeof d;
  d.abi_version = DFX_ABI_VERSION;
  const auto dims = y0.dimensions();
  d.n_traj = dims[0];
  d.dim = static_cast<int32_t>(dims.size() > 1 ? dims[1] : 1);
  d.dtype = DT == ffi::F64 ? DFX_F64 : DFX_F32;
  d.field_id = field_id; d.solver_id = solver_id; d.controller = controller; d.levy_area = levy_area;
  d.field_params = params.begin(); d.n_field_params = static_cast<int32_t>(params.size());
  d.y0 = y0.untyped_data();
  d.t0 = t0; d.t1 = t1; d.dt0 = dt0;
  d.rtol = rtol; d.atol = atol; d.pcoeff = pcoeff; d.icoeff = icoeff; d.dcoeff = dcoeff; d.safety = safety;
  d.factormin = factormin; d.factormax = factormax; d.dtmin = dtmin; d.dtmax = dtmax; d.force_dtmin = force_dtmin;
  d.error_order = error_order;
  d.save_t0 = save_t0; d.save_t1 = save_t1; d.save_steps = save_steps; d.max_steps = max_steps;
  d.n_save_ts = static_cast<int32_t>(save_ts.element_count());
  d.save_ts = d.n_save_ts ? save_ts.untyped_data() : nullptr;
  d.ts_out = ts_out->untyped_data(); d.ys_out = ys_out->untyped_data();
  d.stats = stats->typed_data(); d.result = result->typed_data(); d.y_final = y_final->untyped_data();
  d.bm_keys = keys.element_count() ? keys.typed_data() : nullptr;
  d.bm_t0 = bm_t0; d.bm_t1 = bm_t1; d.bm_tol = bm_tol; d.threefry_partitionable = partitionable;
  const int rc = dfx_ensemble_solve(&d, static_cast<void *>(stream));
  if (rc != DFX_OK) return ffi::Error(ffi::ErrorCode::kInvalidArgument, std::string(dfx_last_error()));
  return ffi::Error::Success();
}

#define DFX_BIND()                                                                                         \
  ffi::Ffi::Bind()                                                                                         \
      .Ctx<ffi::PlatformStream<cudaStream_t>>()                                                            \
      .Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<DT>>().Arg<ffi::Buffer<ffi::U32>>()  \
      .Arg<ffi::Buffer<ffi::F64>>()                                                                        \
      .Ret<ffi::Buffer<DT>>().Ret<ffi::Buffer<DT>>().Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::S32>>() \
      .Ret<ffi::Buffer<DT>>()                                                                              \
      .Attr<int32_t>("field_id").Attr<int32_t>("solver_id").Attr<int32_t>("controller").Attr<int32_t>("levy_area") \
      .Attr<double>("t0").Attr<double>("t1").Attr<double>("dt0").Attr<double>("rtol").Attr<double>("atol")  \
      .Attr<double>("pcoeff").Attr<double>("icoeff").Attr<double>("dcoeff").Attr<double>("safety")         \
      .Attr<double>("factormin").Attr<double>("factormax").Attr<double>("dtmin").Attr<double>("dtmax")      \
      .Attr<int32_t>("force_dtmin").Attr<double>("error_order").Attr<int32_t>("save_t0")                   \
      .Attr<int32_t>("save_t1").Attr<int32_t>("save_steps").Attr<int32_t>("max_steps").Attr<double>("bm_t0") \
      .Attr<double>("bm_t1").Attr<double>("bm_tol").Attr<int32_t>("partitionable")                         \
      .Attr<ffi::Span<const double>>("params")

namespace {
constexpr ffi::DataType kF64 = ffi::F64, kF32 = ffi::F32;
}
#define DT kF64
XLA_FFI_DEFINE_HANDLER_SYMBOL(DfxEnsembleSolveF64, EnsembleSolve<ffi::F64>, DFX_BIND());
#undef DT
#define DT kF32
XLA_FFI_DEFINE_HANDLER_SYMBOL(DfxEnsembleSolveF32, EnsembleSolve<ffi::F32>, DFX_BIND());
#undef DT
