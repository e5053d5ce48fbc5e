import sys
sys.path.insert(0, ".")
import numpy as np
from stroemung_b200 import presets
from stroemung_b200.simulation import Simulation, SOR_RED_BLACK, SOR_REFERENCE_ORDER
from tests.util import unfinalized
shape = (34, 18)
g = presets.simple_inflow(shape)
warm = Simulation.try_from(unfinalized(shape[0], shape[1], g["kind"], g["bu"], g["bv"]))
for t in range(250):
    r = warm.run_simulation_tick()
    if t % 25 == 0: print(t, r)
print("warm last", warm.run_simulation_tick(), warm.initial_norm_squared)
for eps in (1e-3, 1e-6):
    unf = unfinalized(shape[0], shape[1], g["kind"], g["bu"], g["bv"], p=warm.grid.pressure,
                      u=warm.grid.u, v=warm.grid.v, sor_absolute_epsilon=eps,
                      max_iterations=5000, initial_norm_squared=0.0)
    for mode in (SOR_REFERENCE_ORDER, SOR_RED_BLACK):
        s = Simulation.try_from(unf, sor_mode=mode, temporal_block=2)
        print(eps, mode, [s.run_simulation_tick() for _ in range(4)])
