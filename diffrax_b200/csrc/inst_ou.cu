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
// matrix-valued additive diffusion: ControlTerm(lambda t, y, args: G[d, m], VirtualBrownianTree(shape=(m,)))
using M22 = ::dfx::OuMatrixField<2, 2>;
using M32 = ::dfx::OuMatrixField<3, 2>;
using M23 = ::dfx::OuMatrixField<2, 3>;
DFX_REGISTER(double, M22, ::dfx::EulerSolver, 1)
DFX_REGISTER(double, M22, ::dfx::Heun, 1)
DFX_REGISTER(double, M22, ::dfx::SharkSolver, 2)
DFX_REGISTER(float, M22, ::dfx::Heun, 1)
DFX_REGISTER(float, M22, ::dfx::SharkSolver, 2)
DFX_REGISTER(double, M32, ::dfx::Heun, 1)
DFX_REGISTER(double, M32, ::dfx::SharkSolver, 2)
DFX_REGISTER(float, M32, ::dfx::Heun, 1)
DFX_REGISTER(double, M23, ::dfx::Heun, 1)
DFX_REGISTER(double, M23, ::dfx::SharkSolver, 2)
DFX_REGISTER(float, M23, ::dfx::SharkSolver, 2)
}  // namespace
