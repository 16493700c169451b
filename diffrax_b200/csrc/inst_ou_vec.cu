// inst_ou_vec.cu - Ornstein-Uhlenbeck with a vector Brownian motion (shape=(D,), diagonal diffusion).
#include "launch.cuh"
namespace {
// vector Brownian motion, shape=(D,), diagonal diffusion: D = 2, 3
using F2 = ::dfx::OuDiagField<2>;
using F3 = ::dfx::OuDiagField<3>;
DFX_REGISTER(double, F2, ::dfx::EulerSolver, 1)
DFX_REGISTER(double, F2, ::dfx::Heun, 1)
DFX_REGISTER(double, F2, ::dfx::SharkSolver, 2)
DFX_REGISTER(float, F2, ::dfx::Heun, 1)
DFX_REGISTER(float, F2, ::dfx::SharkSolver, 2)
DFX_REGISTER(double, F3, ::dfx::Heun, 1)
DFX_REGISTER(double, F3, ::dfx::SharkSolver, 2)
DFX_REGISTER(float, F3, ::dfx::Heun, 1)
DFX_REGISTER(float, F3, ::dfx::SharkSolver, 2)
}  // namespace
