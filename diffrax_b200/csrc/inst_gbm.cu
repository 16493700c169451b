// inst_gbm.cu - geometric Brownian motion (state-dependent diagonal diffusion): Euler / Heun with a BrownianIncrement tree,
// and adaptive stepping by HalfSolver(Heun).
#include "launch.cuh"
namespace {
using G1 = ::dfx::GbmField<1>;
using G2 = ::dfx::GbmField<2>;
DFX_REGISTER(double, G1, ::dfx::EulerSolver, 1)
DFX_REGISTER(double, G1, ::dfx::Heun, 1)
DFX_REGISTER(float, G1, ::dfx::EulerSolver, 1)
DFX_REGISTER(float, G1, ::dfx::Heun, 1)
DFX_REGISTER(double, G1, ::dfx::HalfOf<::dfx::Heun>, 1)
DFX_REGISTER(double, G2, ::dfx::Heun, 1)
DFX_REGISTER(float, G2, ::dfx::Heun, 1)
}  // namespace
