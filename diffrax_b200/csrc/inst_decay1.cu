// inst_decay1.cu - linear decay, state dimension 1 (one TU per dimension so nvcc runs in parallel)
#include "launch.cuh"
namespace {
using F = ::dfx::DecayField<1>;
DFX_REGISTER_ODE_FIELD(F)
}  // namespace
