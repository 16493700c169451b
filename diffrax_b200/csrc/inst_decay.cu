// inst_decay.cu - kernel instantiations + registry entries (one TU per field so nvcc runs in parallel)
#include "launch.cuh"
namespace {
using F0 = ::dfx::DecayField<1>;
DFX_REGISTER_ODE_FIELD(F0)
using F1 = ::dfx::DecayField<2>;
DFX_REGISTER_ODE_FIELD(F1)
using F2 = ::dfx::DecayField<3>;
DFX_REGISTER_ODE_FIELD(F2)
}  // namespace
