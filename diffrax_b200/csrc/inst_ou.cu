// inst_ou.cu - Ornstein-Uhlenbeck: ODE drift-only kernels plus the SDE kernels
// (Euler / Heun with BrownianIncrement or SpaceTimeLevyArea trees, ShARK with SpaceTimeLevyArea).
#include "launch.cuh"
namespace {
using F = ::dfx::OuField;
DFX_REGISTER_ODE_FIELD(F)
DFX_REGISTER(double, F, ::dfx::EulerSolver, 1)
DFX_REGISTER(double, F, ::dfx::Heun, 1)
DFX_REGISTER(double, F, ::dfx::EulerSolver, 2)
DFX_REGISTER(double, F, ::dfx::Heun, 2)
DFX_REGISTER(double, F, ::dfx::SharkSolver, 2)
DFX_REGISTER(float, F, ::dfx::EulerSolver, 1)
DFX_REGISTER(float, F, ::dfx::Heun, 1)
DFX_REGISTER(float, F, ::dfx::EulerSolver, 2)
DFX_REGISTER(float, F, ::dfx::Heun, 2)
DFX_REGISTER(float, F, ::dfx::SharkSolver, 2)
}  // namespace
