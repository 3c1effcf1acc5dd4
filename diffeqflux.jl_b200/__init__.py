"""diffeqflux.jl_b200 -- B200-native (sm_100a) NeuralODE solve + adjoint behind DiffEqFlux's layer API.

The directory name is not a legal Python identifier; import it through the repo-root shim:

    import diffeqflux_b200 as dfx
    node = dfx.NeuralODE(dfx.Chain(dfx.Dense(2, 50, dfx.tanh), dfx.Dense(50, 2)), (0.0, 1.5), dfx.Tsit5(), saveat=ts)
"""
from . import _native
from ._native import B200Error, build
from .layers import (AdamState, Array, AugmentedNDELayer, BacksolveAdjoint, Chain, Dense, DimMover, FFJORD,
                     EnsembleProblem, EnsembleSerial, EnsembleThreads, FFJORDDistribution, GaussAdjoint, InterpolatingAdjoint, QuadratureAdjoint, NeuralODE, ODEProblem, ODESolution, ReturnCode, Tsit5,
                     UnsupportedError, ZygoteVJP, cube, group_ranges, identity, multiple_shoot, relu, setup, sigmoid,
                     softplus, successful_retcode, tanh, train_step)

__all__ = ["EnsembleProblem", "EnsembleSerial", "EnsembleThreads", "GaussAdjoint", "QuadratureAdjoint", "AdamState", "DimMover", "FFJORDDistribution", "train_step", "Array", "AugmentedNDELayer", "BacksolveAdjoint", "Chain", "Dense", "FFJORD", "InterpolatingAdjoint",
           "NeuralODE", "ODEProblem", "ODESolution", "ReturnCode", "Tsit5", "UnsupportedError", "ZygoteVJP", "cube",
           "group_ranges", "identity", "multiple_shoot", "relu", "setup", "sigmoid", "softplus",
           "successful_retcode", "tanh", "B200Error", "build"]
