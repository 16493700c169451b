// inst_decay3.cu - linear decay, state dimension 3 (one TU per dimension so nvcc runs in parallel)
#include "launch.cuh"
namespace {
using F = ::dfx::DecayField<3>;
DFX_REGISTER_ODE_FIELD(F)
}  // namespace
