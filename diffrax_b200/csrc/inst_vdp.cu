// inst_vdp.cu - van der Pol oscillator: kernel instantiations + registry entries
#include "launch.cuh"
namespace {
using F1 = ::dfx::VdpField;
DFX_REGISTER_ODE_FIELD(F1)
}  // namespace
