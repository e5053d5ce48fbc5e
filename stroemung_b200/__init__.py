"""stroemung_b200 -- B200-native (sm_100a) implementation of stroemung's per-timestep solver.

The compute path is hand-written CUDA behind a C ABI (include/stroemung_b200.h,
built into stroemung_b200/libstroemung_b200.so).  This Python package is the
host-side mirror of the reference's `Simulation` / `SimulationGrid` API
(/root/reference/src/simulation.rs, src/grid/mod.rs) over that ABI via ctypes.
There is no CPU fallback: importing `stroemung_b200.simulation` fails loudly if
the shared library is missing.
"""
__version__ = "0.1.0"
