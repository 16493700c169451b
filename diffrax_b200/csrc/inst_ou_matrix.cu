// inst_ou_matrix.cu - Ornstein-Uhlenbeck drift with a constant [d, m] diffusion matrix (general ControlTerm).
#include "launch.cuh"
namespace {
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
