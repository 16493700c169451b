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
// adaptive SDE stepping by step doubling (docs/usage/getting-started.md:102-110); HalfSolver(Euler()) on an SDE is
// refused by the reference (_integrate.py:1143-1149), so it has no kernel
DFX_REGISTER(double, F, ::dfx::HalfOf<::dfx::Heun>, 1)
DFX_REGISTER(double, F, ::dfx::HalfOf<::dfx::SharkSolver>, 2)
DFX_REGISTER(float, F, ::dfx::HalfOf<::dfx::Heun>, 1)
DFX_REGISTER(float, F, ::dfx::HalfOf<::dfx::SharkSolver>, 2)
}  // namespace
