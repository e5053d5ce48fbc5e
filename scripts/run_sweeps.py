"""Small driver for profiling: build the C5 channel at NX x NY and run a few SOR passes / ticks."""
import argparse
import sys

sys.path.insert(0, ".")
from bench import workload  # noqa: E402
from stroemung_b200.simulation import SOR_RED_BLACK, SOR_REFERENCE_ORDER, Simulation  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, nargs=2, default=(8192, 8192))
ap.add_argument("--workload", default="c5")
ap.add_argument("--mode", default="rb")
ap.add_argument("--tblock", type=int, default=2)
ap.add_argument("--sweeps", type=int, default=8)
ap.add_argument("--ticks", type=int, default=0)
a = ap.parse_args()
wl = workload(a.workload, 1, tuple(a.size))
sim = Simulation.from_preset(wl["preset"], wl["size"], wl["cell_size"], wl["delt"], wl["gamma"],
                             wl["reynolds"], wl["eps"], wl["max_iterations"], wl["omega"],
                             preset_args=wl["preset_args"],
                             sor_mode=SOR_RED_BLACK if a.mode == "rb" else SOR_REFERENCE_ORDER,
                             temporal_block=a.tblock)
if a.ticks:
    print(sim.run_ticks(a.ticks))
if a.sweeps:
    norms = sim.sor_sweeps(a.sweeps)
    print("sor ms", sim.last_sor_ms, "last norm", norms[-1])
