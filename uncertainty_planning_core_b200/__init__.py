"""B200-native particle hot path of uncertainty_planning_core.

Batched noisy forward simulation with sphere-vs-SDF collision and contact resolution
(SimulatorInterface::ForwardSimulateRobots) and outcome clustering (ClusterParticles), as hand-written
sm_100a CUDA kernels behind a C ABI (include/upc_b200.h).  There is no CPU fallback: importing the
simulator without the built library raises ImportError.
"""
from . import abi  # noqa: F401
from .simulator import B200Simulator, OutcomeClustering  # noqa: F401

__all__ = ["abi", "B200Simulator", "OutcomeClustering"]
