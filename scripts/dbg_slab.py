"""Debug helper: where do slab runs first differ from the single-GPU run?"""
import sys
import numpy as np
sys.path.insert(0, ".")
from stroemung_b200 import multi
from stroemung_b200.simulation import SOR_RED_BLACK

preset, size, args = sys.argv[1], (int(sys.argv[2]), int(sys.argv[3])), ()
world, T, max_it, ticks = int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7])


def run(group):
    sim = multi.from_preset(group, preset, size, (0.1, 0.2), 0.005, 0.9, 100.0, 1e-3, max_it, 1.7,
                            preset_args=args, temporal_block=T, device=0)
    snaps = []

    def snap(tag):
        d = {k: multi.gather_field(group, getattr(sim.grid, k)) for k in ("pressure", "u", "v")}
        d.update({k: multi.gather_field(group, getattr(sim, k)) for k in ("f", "g", "rhs")})
        d["edge"] = multi.gather_field(group, sim.grid.edge_type)
        snaps.append((tag, d))
    snap("created")
    for t in range(ticks):
        if len(sys.argv) > 8:
            r = sim.run_simulation_tick(); snap(f"t{t} tick {r}")
            continue
        sim.grid.set_boundary_u_and_v(); snap(f"t{t} bc")
        sim.calculate_f_and_g(); snap(f"t{t} fg")
        sim.calculate_rhs(); snap(f"t{t} rhs")
        r = sim.solve_sor(); snap(f"t{t} sor {r}")
        sim.set_u_and_v(); snap(f"t{t} adapt")
    group.barrier()
    sim.close()
    return snaps


ref = multi.run_threads(1, run)[0]
got = multi.run_threads(world, run)[0]
for (tag, a), (tag2, b) in zip(ref, got):
    line = [f"{tag:28s}|{tag2:28s}"]
    for k in a:
        bad = np.nonzero(a[k] != b[k])
        if bad[0].size:
            line.append(f"{k}: {bad[0].size} bad rows {bad[0].min()}..{bad[0].max()} cols {bad[1].min()}..{bad[1].max()}")
    print(" ".join(line))
