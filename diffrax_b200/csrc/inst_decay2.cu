// inst_decay2.cu - linear decay, state dimension 2 (one TU per dimension so nvcc runs in parallel)
#include "launch.cuh"
namespace {
using F = ::dfx::DecayField<2>;
DFX_REGISTER_ODE_FIELD(F)
}  // namespace
