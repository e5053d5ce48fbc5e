"""Row-slab (multi-GPU) path against the single-GPU path and the oracle.

The slab handles are driven by threads of this one process (stroemung_b200.multi.run_threads)
so that the whole exchange machinery -- P2P halo stores from the red-black pass, the put
kernel, the mailbox all-gather, the collective classification -- runs even on a box with a
single B200: slab r is placed on device r % device_count.  With >= 2 GPUs the same tests
cross NVLink.  (The one-process-per-GPU launch through CUDA IPC is covered by
tests/slab_worker.py under torchrun, see test_torchrun_two_ranks.)

Bar: p, u, v, f, g, rhs bit-identical to the single-GPU red-black run (the slab edges must
be invisible); sweep counts equal; residual norms within 1e-12 relative (per-slab partial
sums are added in rank order instead of tile order).
"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from stroemung_b200 import multi, presets
from stroemung_b200.simulation import SOR_RED_BLACK, BoundaryTooThinError, Simulation
from tests.util import DEFAULTS, assert_bits_equal, random_fields, random_mask, unfinalized

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def n_devices():
    import torch
    return torch.cuda.device_count()


def close(a, b, rtol=1e-12):
    return a == b or abs(a - b) <= rtol * max(abs(a), abs(b))


PRM = dict(cell_size=(0.1, 0.2), delt=0.005, gamma=0.9, reynolds=100.0,
           sor_absolute_epsilon=1e-3, max_iterations=100, omega=1.7)


def preset_run(group, preset, size, preset_args, T, ticks, max_it):
    nd = n_devices()
    sim = multi.from_preset(group, preset, size, PRM["cell_size"], PRM["delt"], PRM["gamma"],
                            PRM["reynolds"], PRM["sor_absolute_epsilon"], max_it, PRM["omega"],
                            preset_args=preset_args, temporal_block=T,
                            device=group.rank % nd)
    res = [sim.run_simulation_tick() for _ in range(ticks)]
    out = {k: multi.gather_field(group, getattr(sim.grid, k)) for k in ("pressure", "u", "v")}
    out.update({k: multi.gather_field(group, getattr(sim, k)) for k in ("f", "g", "rhs")})
    out["res"] = res
    out["speed_range"] = sim.grid.speed_range
    out["pressure_range"] = sim.grid.pressure_range
    out["fluid_cells"] = sim.grid.boundaries.fluid_cells
    out["initial_norm"] = sim.initial_norm_squared
    group.barrier()
    sim.close()
    return out


def compare_runs(ref, got, what):
    for k in ("pressure", "u", "v", "f", "g", "rhs"):
        assert_bits_equal(got[k], ref[k], f"{what}: {k}")
    assert [r[0] for r in got["res"]] == [r[0] for r in ref["res"]], what
    for (_, a), (_, b) in zip(got["res"], ref["res"]):
        assert close(a, b), (what, a, b)
    assert got["speed_range"] == ref["speed_range"]
    assert got["pressure_range"] == ref["pressure_range"]
    assert got["fluid_cells"] == ref["fluid_cells"]
    assert close(got["initial_norm"], ref["initial_norm"])


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("preset,size,args,T,ticks,max_it", [
    ("simple_inflow", (64, 40), (), 2, 4, 30),
    ("obstacle", (100, 20), (), 4, 3, 25),
    ("obstacle", (100, 20), (), 1, 2, 10),
    ("channel_circle", (150, 70), (31, 35, 9.0), 3, 3, 20),   # circle straddles a slab edge
    ("backward_step", (96, 48), (33, 24), 4, 3, 17),
    ("cavity", (48, 130), (1.0,), 2, 3, 12),                   # two tile columns
])
def test_slabs_match_single_gpu(world, preset, size, args, T, ticks, max_it):
    ref = multi.run_threads(1, lambda g: preset_run(g, preset, size, args, T, ticks, max_it))[0]
    got = multi.run_threads(world, lambda g: preset_run(g, preset, size, args, T, ticks, max_it))
    for r in range(world):
        compare_runs(ref, got[r], f"{preset} world {world} rank {r}")


def test_slabs_from_host_arrays_random_state():
    """try_from with per-slab host arrays: flag halos, merged velocity tables (interior
    Inflow blocks next to slab edges), random p / u / v."""
    nx, ny, world = 90, 37, 3
    kind, bu, bv = random_mask(nx, ny, 41, n_blocks=8)
    p, u, v = random_fields(nx, ny, 41)
    over = dict(max_iterations=23)
    full = unfinalized(nx, ny, kind, bu, bv, p=p, u=u, v=v, **over)

    def one(group):
        if group.world == 1:
            unf = full
        else:
            xb, xe = multi.slab_range(nx, group.rank, group.world)
            unf = unfinalized(nx, ny, kind[xb:xe], bu[xb:xe], bv[xb:xe], p=p[xb:xe], u=u[xb:xe],
                              v=v[xb:xe], **over)
        sim = multi.try_from(group, unf, sor_mode=SOR_RED_BLACK, temporal_block=3,
                             device=group.rank % n_devices())
        res = [sim.run_simulation_tick() for _ in range(3)]
        out = {k: multi.gather_field(group, getattr(sim.grid, k))
               for k in ("pressure", "u", "v", "edge_type", "cell_type")}
        out.update({k: multi.gather_field(group, getattr(sim, k)) for k in ("f", "g", "rhs")})
        out.update(res=res, speed_range=sim.grid.speed_range,
                   pressure_range=sim.grid.pressure_range,
                   fluid_cells=sim.grid.boundaries.fluid_cells,
                   initial_norm=sim.initial_norm_squared)
        # new fields on every slab, halos refreshed collectively, one more tick
        xb, xe = multi.slab_range(nx, group.rank, group.world)
        sim.grid.u = v[xb:xe] if group.world > 1 else v
        sim.grid.pressure = u[xb:xe] if group.world > 1 else u
        if group.world > 1:
            sim.slab_sync_halos()
        out["res2"] = sim.run_simulation_tick()
        out["u2"] = multi.gather_field(group, sim.grid.u)
        out["p2"] = multi.gather_field(group, sim.grid.pressure)
        group.barrier()
        sim.close()
        return out

    ref = multi.run_threads(1, one)[0]
    got = multi.run_threads(world, one)
    for r in range(world):
        compare_runs(ref, got[r], f"random rank {r}")
        assert np.array_equal(got[r]["edge_type"], ref["edge_type"])
        assert np.array_equal(got[r]["cell_type"], ref["cell_type"])
        assert got[r]["res2"][0] == ref["res2"][0] and close(got[r]["res2"][1], ref["res2"][1])
        assert_bits_equal(got[r]["u2"], ref["u2"], "u after upload + sync")
        assert_bits_equal(got[r]["p2"], ref["p2"], "p after upload + sync")


def test_slab_sweep_norms_and_oracle():
    """sb_sor_sweeps on slabs: all ranks return the same norms, equal to the oracle's
    red-black restatement within 1e-12; the field is bit-exact."""
    from tests.util import oracle_from
    from oracle import pyoracle as po
    nx, ny, world, T, n = 72, 50, 2, 4, 11
    kind, bu, bv = random_mask(nx, ny, 77)
    p, u, v = random_fields(nx, ny, 77)
    full = unfinalized(nx, ny, kind, bu, bv, p=p, u=u, v=v)
    o = oracle_from(full, sor_mode=po.SOR_RED_BLACK)
    onorms = []
    for _ in range(n):
        o.sor_sweep()
        onorms.append(o.calculate_norm_squared())

    def one(group):
        xb, xe = multi.slab_range(nx, group.rank, group.world)
        unf = unfinalized(nx, ny, kind[xb:xe], bu[xb:xe], bv[xb:xe], p=p[xb:xe], u=u[xb:xe],
                          v=v[xb:xe])
        sim = multi.try_from(group, unf, sor_mode=SOR_RED_BLACK, temporal_block=T,
                             device=group.rank % n_devices())
        norms = sim.sor_sweeps(n)
        pf = multi.gather_field(group, sim.grid.pressure)
        group.barrier()
        sim.close()
        return norms, pf

    got = multi.run_threads(world, one)
    for norms, pf in got:
        assert np.array_equal(norms, got[0][0])
        for a, b in zip(norms, onorms):
            assert close(a, b), (a, b)
        assert_bits_equal(pf, o.p, "p after slab sweeps vs oracle")


def test_slab_thin_boundary_reported_on_every_rank():
    """A one-cell-thick wall inside slab 1: all ranks must raise the same error cell."""
    nx, ny, world = 60, 16, 2
    g = presets.simple_inflow((nx, ny))
    kind = g["kind"].copy()
    kind[40:44, 8] = 1          # 1-wide in y: fluid north and south -> too thin
    bu, bv = g["bu"], g["bv"]

    def one(group):
        xb, xe = multi.slab_range(nx, group.rank, group.world)
        unf = unfinalized(nx, ny, kind[xb:xe], bu[xb:xe], bv[xb:xe])
        try:
            multi.try_from(group, unf, sor_mode=SOR_RED_BLACK, device=group.rank % n_devices())
        except BoundaryTooThinError as e:
            return e.xy
        return None

    got = multi.run_threads(world, one)
    assert got == [(40, 8)] * world


def test_torchrun_two_ranks():
    """One process per GPU through CUDA IPC (the production launch).  Needs 2 GPUs."""
    if n_devices() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29541",
           str(ROOT / "tests" / "slab_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, PYTHONPATH=str(ROOT)))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "slab_worker ok" in r.stdout
