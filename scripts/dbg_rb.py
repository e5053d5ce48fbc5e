import sys
sys.path.insert(0, ".")
import numpy as np
from stroemung_b200 import presets
from stroemung_b200.simulation import Simulation, SOR_RED_BLACK
from tests.util import unfinalized
size = tuple(int(a) for a in sys.argv[1:3]) if len(sys.argv) > 2 else (34, 18)
g = presets.simple_inflow(size)
unf = unfinalized(size[0], size[1], g["kind"], g["bu"], g["bv"])
sim = Simulation.try_from(unf, sor_mode=SOR_RED_BLACK, temporal_block=1)
print("created; initial norm", sim.initial_norm_squared)
print(sim.sor_sweeps(3))
