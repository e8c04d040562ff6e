"""b200dde — host-side mirror of DelayDiffEq.jl's MethodOfSteps ensemble API over libb200dde.so (sm_100a)."""
from ._lib import B200DDEError, Config, MODELS, RETCODES, STAT_NAMES, LIB_PATH, exported_symbols, lib  # noqa: F401
from .api import (BS3, DDEProblem, Discontinuity, NLAnderson, DDESolution, DEStats, EnsembleB200, EnsembleProblem, EnsembleSolution, MethodOfSteps,  # noqa: F401
                  NLFunctional, Rosenbrock23, Solver, Tsit5, UserModel, saveat_grid, solve, DiscreteCallback, PresetTimeCallback,
                  PeriodicCallback, ContinuousCallback)
