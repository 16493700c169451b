// inst_osc.cu - forced oscillator: kernel instantiations + registry entries (one TU per field so nvcc runs in parallel)
#include "launch.cuh"
namespace {
using F0 = ::dfx::ForcedOscField;
DFX_REGISTER_ODE_FIELD(F0)
}  // namespace
