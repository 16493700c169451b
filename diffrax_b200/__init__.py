"""diffrax_b200 - B200-native ensemble integrator behind Diffrax's diffeqsolve interface.

Scope: ``diffeqsolve`` over a batch of independent trajectories with explicit
Runge-Kutta solvers, PIDController / ConstantStepSize, SaveAt (t0/t1/ts/steps/dense) and
VirtualBrownianTree-driven Heun / ShARK / Euler SDE solves.  All numerics run in hand-written
sm_100a CUDA kernels (csrc/) behind the C ABI of include/diffrax_b200.h.
"""
from . import fields, random
from ._api import (RESULTS, AffineEvent, Event, Newton, SteadyStateEvent, is_event, is_okay, steady_state_event, Bosh3, BrownianIncrement, ClipStepSizeController, ConstantStepSize, ControlTerm, DenseInterpolation,
                   Dopri5, Dopri8, Euler, HalfSolver, Heun, Midpoint, MultiTerm, ODETerm, PIDController, Ralston, SaveAt,
                   ShARK, Solution, SpaceTimeLevyArea, SubSaveAt, Tsit5, VirtualBrownianTree, diffeqsolve, is_successful, prepare, EnsembleSolve)

from ._dist import ShardedSolution, prepare_sharded, shard_range, sharded_diffeqsolve  # noqa: E402

__all__ = [
    "ShardedSolution", "prepare_sharded", "shard_range", "sharded_diffeqsolve",
    "RESULTS", "AffineEvent", "Event", "Newton", "SteadyStateEvent", "is_event", "is_okay", "steady_state_event", "Bosh3", "BrownianIncrement", "ClipStepSizeController", "ConstantStepSize", "ControlTerm", "DenseInterpolation", "Dopri5",
    "Dopri8", "Euler", "HalfSolver", "Heun", "Midpoint", "MultiTerm", "ODETerm", "PIDController", "Ralston", "SaveAt", "ShARK",
    "Solution", "SpaceTimeLevyArea", "SubSaveAt", "Tsit5", "VirtualBrownianTree", "diffeqsolve", "is_successful", "prepare",
    "EnsembleSolve", "fields",
    "random",
]
