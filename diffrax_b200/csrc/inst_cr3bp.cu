// inst_cr3bp.cu - kernel instantiations + registry entries (one TU per field so nvcc runs in parallel)
#include "launch.cuh"
namespace {
using F0 = ::dfx::Cr3bpField;
DFX_REGISTER_ODE_FIELD(F0)
}  // namespace
