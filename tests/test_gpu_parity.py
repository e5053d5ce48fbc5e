"""Parity of the CUDA path (through the C ABI) with the oracle and the reference's goldens.

Bar (BASELINE.json north_star):
  * integer work (classification, boundary list): bit-exact;
  * reference-order mode: every field bit-exact per stage and per tick; the residual
    norm within 1e-12 relative (its summation order is the only freedom);
  * performance mode (red-black): bit-exact against the oracle's red-black restatement,
    and within the SOR-eps tolerance of the reference-order solution on converged ticks.
"""
import numpy as np
import pytest

from oracle import pyoracle as po
from stroemung_b200 import math as sbmath
from stroemung_b200 import presets, refjson
from stroemung_b200.simulation import (SOR_RED_BLACK, SOR_REFERENCE_ORDER, BoundaryTooThinError,
                                       Simulation)
from tests.util import (DEFAULTS, assert_bits_equal, oracle_from, random_fields, random_mask,
                        unfinalized)

pytestmark = pytest.mark.gpu


SOR_PATH = "resident-kernels"


@pytest.fixture(autouse=True, params=["resident-kernels", "mid-kernel", "pass-kernels"])
def sor_path_switch(request, monkeypatch):
    """By default a red-black solve takes the one-launch kernels where the grid fits them:
    sor_small.cu (up to 12 288 cells, one SM) and sor_mid.cu (the shared memory of all SMs,
    cooperative launch).  Every test here also runs with the small kernel off (small shapes
    then take the mid kernel on a few CTAs) and with both off, so that all shapes keep
    exercising the tile / streaming pass kernels."""
    global SOR_PATH
    SOR_PATH = request.param
    if request.param != "resident-kernels":
        monkeypatch.setenv("SB_SOR_SMALL", "0")
    if request.param == "pass-kernels":
        monkeypatch.setenv("SB_SOR_MID", "0")
    yield
    SOR_PATH = "resident-kernels"

TICK = "stroemung__simulation__tests__simulation_tick"
NORM_RTOL = 1e-12


def close(a, b, rtol=NORM_RTOL):
    return a == b or abs(a - b) <= rtol * max(abs(a), abs(b))


# ---- known-answer tests on the device: src/math.rs:192-402, src/simulation.rs:449-569 ----
def test_kat_operators_on_device(kat):
    for c in kat["du2dx"]:
        assert sbmath.du2dx(c["u"], c["delx"], c["gamma"]) == c["expected"]
    for c in kat["dv2dy"]:
        assert sbmath.dv2dy(c["v"], c["dely"], c["gamma"]) == c["expected"]
    for c in kat["duvdx"]:
        assert sbmath.duvdx(c["u"], c["v"], c["delx"], c["gamma"]) == c["expected"]
    for c in kat["duvdy"]:
        assert sbmath.duvdy(c["u"], c["v"], c["dely"], c["gamma"]) == c["expected"]
    for c in kat["laplacian"]:
        assert sbmath.laplacian(c["e"], c["delx"], c["dely"]) == c["expected"]
    for name in ("calculate_f", "calculate_g"):
        for c in kat[name]:
            got = getattr(sbmath, name)(c["u"], c["v"], c["delx"], c["dely"], c["delt"],
                                        c["gamma"], c["reynolds"])
            assert got == c["expected"], (name, got, c["expected"])


def test_residual_operator_matches_oracle():
    rng = np.random.default_rng(7)
    for _ in range(20):
        blk = rng.uniform(-3, 3, 9)
        rhs = float(rng.uniform(-5, 5))
        assert sbmath.residual(blk, 0.1, 0.2, rhs) == po.residual(blk, 0.1, 0.2, rhs)


# ---- the reference's simulation_tick test, on the GPU (src/simulation.rs:571-618) --------
def check_sim_snapshot(sim, snap):
    assert_bits_equal(sim.grid.pressure, refjson.array_from_json(snap["grid"]["pressure"]), "p")
    assert_bits_equal(sim.grid.u, refjson.array_from_json(snap["grid"]["u"]), "u")
    assert_bits_equal(sim.grid.v, refjson.array_from_json(snap["grid"]["v"]), "v")
    assert sim.time == snap["time"]
    assert sim.iterations == snap["iterations"]
    assert sim.initial_norm_squared == snap["initial_norm_squared"]
    assert np.array_equal(sim.grid.cell_type,
                          refjson.cells_from_json(snap["grid"]["cell_type"])[0])


def test_simulation_tick_golden(kat, snapshots):
    size = (4, 3)
    unf = unfinalized(4, 3, **{k: presets.simple_inflow(size)[k] for k in ("kind", "bu", "bv")})
    sim = Simulation.try_from(unf)
    it, nrm = sim.run_simulation_tick()
    for suffix, field in (("", sim.f), ("-2", sim.g), ("-3", sim.rhs)):
        assert_bits_equal(field, refjson.array_from_json(snapshots[TICK + suffix]["json"]),
                          "tick1" + suffix)
    check_sim_snapshot(sim, snapshots[TICK + "-4"]["json"])
    a = kat["simulation_tick_asserts"]
    assert it == a[0]["sor_iterations"] and close(nrm, a[0]["norm_squared"])
    for _ in range(100):
        it, nrm = sim.run_simulation_tick()
    assert it == a[1]["sor_iterations"] and close(nrm, a[1]["norm_squared"])
    for suffix, field in (("-5", sim.f), ("-6", sim.g), ("-7", sim.rhs)):
        assert_bits_equal(field, refjson.array_from_json(snapshots[TICK + suffix]["json"]),
                          "tick101" + suffix)
    check_sim_snapshot(sim, snapshots[TICK + "-8"]["json"])
    sim.run_ticks(100)
    check_sim_snapshot(sim, snapshots[TICK + "-9"]["json"])


def test_deserialize_golden(fixtures, snapshots):
    raw = fixtures["src/test_data/small_simulation_with_boundaries.json"]["raw"]
    prm, grid = refjson.simulation_from_json(refjson.loads(raw, quirk_serde_json=True))
    prm["grid"] = grid
    sim = Simulation.try_from(prm)
    snap = snapshots["stroemung__simulation__tests__deserialize-2"]["json"]
    assert close(sim.initial_norm_squared, snap["initial_norm_squared"])
    assert close(sim.initial_norm_squared, 899.9547140394143)
    # NaSt2D-derived state: one tick must match the oracle bit for bit as well
    o = oracle_from(prm)
    assert sim.run_simulation_tick()[0] == o.run_simulation_tick()[0]
    assert_bits_equal(sim.grid.pressure, o.p, "p")
    assert_bits_equal(sim.grid.u, o.u, "u")
    # 5x7 all-fluid grid of the other fixture
    prm, grid = refjson.simulation_from_json(
        fixtures["src/test_data/simple_simulation.json"]["json"])
    prm["grid"] = grid
    sim = Simulation.try_from(prm)
    assert sim.initial_norm_squared == 0.0
    assert sim.grid.boundaries.fluid_cells == 35.0


# ---- boundary classification (integer, bit-exact): src/grid/mod.rs:683-823 --------------
def grid3(cells):
    kind = np.zeros((3, 3), dtype=np.uint8)
    for c in cells:
        kind[c] = 1
    return kind


def test_thin_boundary():
    for cells, first in (([(1, 0), (1, 1), (1, 2)], (1, 0)), ([(0, 1), (1, 1), (2, 1)], (0, 1))):
        kind = grid3(cells)
        with pytest.raises(BoundaryTooThinError) as ei:
            Simulation.try_from(unfinalized(3, 3, kind, None, None))
        with pytest.raises(po.BoundaryTooThin) as eo:
            po.OracleSim(3, 3, kind=kind, **DEFAULTS)
        assert ei.value.xy == tuple(int(x) for x in eo.value.xy) == first
        assert ei.value.kind == 1


def test_rebuild_boundary_list():
    examples = [
        ([(0, 0), (0, 1), (0, 2), (1, 0), (1, 2), (2, 0), (2, 1), (2, 2)],
         [None, "East", None, "South", "North", None, "West", None]),
        ([(0, 0), (0, 2), (2, 0), (2, 2)], ["SouthEast", "NorthEast", "SouthWest", "NorthWest"]),
    ]
    for cells, edges in examples:
        sim = Simulation.try_from(unfinalized(3, 3, grid3(cells), None, None))
        assert sim.grid.boundaries.sorted_boundary_list == list(zip(cells, edges))
        assert sim.grid.boundaries.fluid_cells == 9 - len(cells)


@pytest.mark.parametrize("shape,seed", [((34, 18), 1), ((64, 48), 2), ((257, 129), 3)])
def test_classification_random_masks(shape, seed):
    nx, ny = shape
    kind, bu, bv = random_mask(nx, ny, seed)
    sim = Simulation.try_from(unfinalized(nx, ny, kind, bu, bv))
    o = po.OracleSim(nx, ny, kind=kind, bu=bu, bv=bv, **DEFAULTS)
    idx, edge = sim.grid._boundary_arrays()
    oidx, oedge = o.boundary_list()
    assert np.array_equal(idx, oidx)
    assert np.array_equal(edge, oedge)
    assert sim.grid.boundaries.fluid_cells == o.state().fluid_cells
    assert np.array_equal(sim.grid.cell_type, kind)
    # every edge class occurs in the larger masks
    if nx >= 64:
        assert set(int(e) for e in oedge) == set(range(9))
    # a thin wall somewhere in the middle: same first offender as the oracle
    bad = kind.copy()
    bad[nx // 2, 1:ny - 1] = 1
    bad[nx // 2 + 1, 1:ny - 1] = 0
    bad[nx // 2 - 1, 1:ny - 1] = 0
    with pytest.raises(BoundaryTooThinError) as ei:
        Simulation.try_from(unfinalized(nx, ny, bad, bu, bv))
    with pytest.raises(po.BoundaryTooThin) as eo:
        po.OracleSim(nx, ny, kind=bad, bu=bu, bv=bv, **DEFAULTS)
    assert ei.value.xy == tuple(int(x) for x in eo.value.xy)


# ---- stage-level parity, reference-order mode, random mixed-kind grids -------------------
@pytest.mark.parametrize("shape,seed", [((34, 18), 11), ((100, 20), 12), ((257, 129), 13),
                                        ((300, 70), 14)])
def test_stages_bit_exact(shape, seed):
    nx, ny = shape
    kind, bu, bv = random_mask(nx, ny, seed)
    p, u, v = random_fields(nx, ny, seed)
    unf = unfinalized(nx, ny, kind, bu, bv, p=p, u=u, v=v)
    sim = Simulation.try_from(unf)
    o = oracle_from(unf)
    # construction: F/G, RHS and initial norm of the raw state (no velocity BC first)
    assert_bits_equal(sim.f, o.f, "f0")
    assert_bits_equal(sim.g, o.g, "g0")
    assert_bits_equal(sim.rhs, o.rhs, "rhs0")
    assert close(sim.initial_norm_squared, o.state().initial_norm_squared)
    assert sim.grid.pressure_range == list(o.state().pressure_range)
    assert sim.grid.speed_range == list(o.state().speed_range)
    # velocity BC (sequential in the reference, gather on the GPU)
    sim.grid.set_boundary_u_and_v()
    o.set_boundary_u_and_v()
    assert_bits_equal(sim.grid.u, o.u, "u after BC")
    assert_bits_equal(sim.grid.v, o.v, "v after BC")
    sim.calculate_f_and_g()
    o.calculate_f_and_g()
    assert_bits_equal(sim.f, o.f, "f")
    assert_bits_equal(sim.g, o.g, "g")
    sim.calculate_rhs()
    o.calculate_rhs()
    assert_bits_equal(sim.rhs, o.rhs, "rhs")
    # pressure BC + three lexicographic sweeps, norm after each
    for k in range(3):
        sim.grid.copy_pressure_to_boundaries()
        o.copy_pressure_to_boundaries()
        assert_bits_equal(sim.grid.pressure, o.p, f"p after BC {k}")
        norms = sim.sor_sweeps(1)
        o.sor_sweep()
        assert_bits_equal(sim.grid.pressure, o.p, f"p after sweep {k}")
        assert close(norms[0], o.calculate_norm_squared())
        assert close(sim.calculate_norm_squared(), o.calculate_norm_squared())
    # velocity update incl. the restore quirk and the speed range
    sim.set_u_and_v()
    o.set_u_and_v()
    assert_bits_equal(sim.grid.u, o.u, "u after update")
    assert_bits_equal(sim.grid.v, o.v, "v after update")
    assert sim.grid.speed_range == list(o.state().speed_range)
    sim.grid.calculate_pressure_range()
    o.calculate_pressure_range()
    assert sim.grid.pressure_range == list(o.state().pressure_range)


@pytest.mark.parametrize("preset,shape,ticks", [("obstacle", (100, 20), 12),
                                                ("simple_inflow", (34, 18), 120),
                                                ("obstacle", (130, 66), 4)])
def test_ticks_reference_order(preset, shape, ticks):
    """Whole ticks: config 1 of BASELINE.json (default obstacle preset) and a channel that
    reaches the converged regime (early exits after a few sweeps)."""
    g = getattr(presets, preset)(shape)
    unf = unfinalized(shape[0], shape[1], g["kind"], g["bu"], g["bv"])
    sim = Simulation.try_from(unf)
    o = oracle_from(unf)
    for t in range(ticks):
        it, nrm = sim.run_simulation_tick()
        oit, onrm = o.run_simulation_tick()
        assert it == oit, (t, it, oit)
        assert close(nrm, onrm), (t, nrm, onrm)
    assert_bits_equal(sim.grid.pressure, o.p, "p")
    assert_bits_equal(sim.grid.u, o.u, "u")
    assert_bits_equal(sim.grid.v, o.v, "v")
    assert_bits_equal(sim.f, o.f, "f")
    assert_bits_equal(sim.rhs, o.rhs, "rhs")
    assert sim.time == o.state().time
    assert sim.grid.speed_range == list(o.state().speed_range)
    assert sim.grid.pressure_range == list(o.state().pressure_range)


def test_ticks_random_state_reference_order():
    nx, ny = 96, 40
    kind, bu, bv = random_mask(nx, ny, 21)
    p, u, v = random_fields(nx, ny, 21, scale=0.2)
    unf = unfinalized(nx, ny, kind, bu, bv, p=p, u=u, v=v, max_iterations=17)
    sim = Simulation.try_from(unf)
    o = oracle_from(unf)
    for t in range(5):
        it, nrm = sim.run_simulation_tick()
        oit, onrm = o.run_simulation_tick()
        assert it == oit and close(nrm, onrm), (t, it, oit, nrm, onrm)
        assert_bits_equal(sim.grid.pressure, o.p, f"p tick {t}")
        assert_bits_equal(sim.grid.u, o.u, f"u tick {t}")
        assert_bits_equal(sim.grid.v, o.v, f"v tick {t}")


# ---- performance mode: red-black, temporally blocked -------------------------------------
@pytest.mark.parametrize("T", [1, 2, 3, 4])
@pytest.mark.parametrize("shape,seed", [((34, 18), 31), ((100, 20), 32), ((257, 129), 33),
                                        ((150, 300), 34)])
def test_red_black_sweeps_match_oracle(shape, seed, T):
    nx, ny = shape
    kind, bu, bv = random_mask(nx, ny, seed)
    p, u, v = random_fields(nx, ny, seed)
    unf = unfinalized(nx, ny, kind, bu, bv, p=p, u=u, v=v)
    sim = Simulation.try_from(unf, sor_mode=SOR_RED_BLACK, temporal_block=T)
    o = oracle_from(unf, sor_mode=po.SOR_RED_BLACK)
    assert close(sim.initial_norm_squared, o.state().initial_norm_squared)
    n = 7  # not a multiple of T for T in 2..4: exercises the shortened last pass
    norms = sim.sor_sweeps(n)
    for k in range(n):
        o.sor_sweep()
        assert close(norms[k], o.calculate_norm_squared()), (k, norms[k])
    assert_bits_equal(sim.grid.pressure, o.p, "p after red-black sweeps")
    assert close(sim.calculate_norm_squared(), o.calculate_norm_squared())


@pytest.mark.parametrize("T", [1, 2, 3, 4])
@pytest.mark.parametrize("shape,blocks,seed", [((200, 300), 0, 41), ((330, 420), 3, 42),
                                               ((130, 700), 1, 43)])
def test_red_black_streaming_regions_match_oracle(shape, blocks, seed, T):
    """Grids large enough for all-fluid regions: those go through the streaming kernel
    (sor_rb_stream.cu), the rest through the tile kernel; together bit-exact vs the oracle,
    including the shortened last pass (7 sweeps) and the norm of every sweep."""
    nx, ny = shape
    kind, bu, bv = random_mask(nx, ny, seed, n_blocks=blocks)
    p, u, v = random_fields(nx, ny, seed)
    unf = unfinalized(nx, ny, kind, bu, bv, p=p, u=u, v=v)
    sim = Simulation.try_from(unf, sor_mode=SOR_RED_BLACK, temporal_block=T)
    o = oracle_from(unf, sor_mode=po.SOR_RED_BLACK)
    n = 7
    norms = sim.sor_sweeps(n)
    slow, items = sim.rb_plan
    if SOR_PATH == "pass-kernels":
        assert items > 0 or blocks > 1, (slow, items)
    for k in range(n):
        o.sor_sweep()
        assert close(norms[k], o.calculate_norm_squared()), (k, norms[k])
    assert_bits_equal(sim.grid.pressure, o.p, "p after red-black sweeps")
    for t in range(2):   # full ticks on top (exit test, redo passes, velocity update)
        it, nrm = sim.run_simulation_tick()
        oit, onrm = o.run_simulation_tick()
        assert it == oit and close(nrm, onrm), (t, it, oit, nrm, onrm)
    assert_bits_equal(sim.grid.pressure, o.p, "p after ticks")
    assert_bits_equal(sim.grid.u, o.u, "u after ticks")


@pytest.mark.parametrize("T", [1, 2, 4])
def test_red_black_ticks_match_oracle(T):
    shape = (100, 20)
    g = presets.obstacle(shape)
    unf = unfinalized(shape[0], shape[1], g["kind"], g["bu"], g["bv"])
    sim = Simulation.try_from(unf, sor_mode=SOR_RED_BLACK, temporal_block=T)
    o = oracle_from(unf, sor_mode=po.SOR_RED_BLACK)
    for t in range(6):
        it, nrm = sim.run_simulation_tick()
        oit, onrm = o.run_simulation_tick()
        assert it == oit and close(nrm, onrm), (t, it, oit, nrm, onrm)
    assert_bits_equal(sim.grid.pressure, o.p, "p")
    assert_bits_equal(sim.grid.u, o.u, "u")
    assert_bits_equal(sim.grid.v, o.v, "v")


@pytest.mark.parametrize("T", [1, 2, 3, 4])
def test_red_black_early_exit_inside_block(T):
    """Converged regime: the exit test fires after a few sweeps, often inside a temporal
    block; the pass is then redone with the exact count and must agree with the oracle."""
    shape = (34, 18)
    g = presets.simple_inflow(shape)
    unf = unfinalized(shape[0], shape[1], g["kind"], g["bu"], g["bv"])
    sim = Simulation.try_from(unf, sor_mode=SOR_RED_BLACK, temporal_block=T)
    o = oracle_from(unf, sor_mode=po.SOR_RED_BLACK)
    seen = set()
    for t in range(260):
        it, nrm = sim.run_simulation_tick()
        oit, onrm = o.run_simulation_tick()
        assert it == oit and close(nrm, onrm), (t, it, oit, nrm, onrm)
        seen.add(it)
    assert len(seen) > 2 and min(seen) < 100, seen
    assert_bits_equal(sim.grid.pressure, o.p, "p")
    assert_bits_equal(sim.grid.u, o.u, "u")


@pytest.mark.parametrize("eps", [1e-3, 1e-6])
def test_red_black_vs_reference_order_converged(eps):
    """Performance mode against the REFERENCE ordering (SURVEY.md 8a A6 protocol): compare
    on converged ticks, pressure up to its free constant; tolerance: the SOR epsilon
    (larger grids, an obstacle and tighter eps: tests/test_gpu_scale.py).

    The comparison needs a solvable pressure problem: in the inflow/outflow channel presets
    the discrete Neumann problem is slightly inconsistent, the residual norm has a floor
    (measured on the B200: ~1e-7 for the lexicographic order, ~4e-6 for red-black at tick
    250 of the 34x18 channel) and "SOR to eps" is then decided by the cap, not by eps.  A
    closed cavity with a moving lid has no net flux, both orderings converge every tick."""
    shape = (34, 34)
    g = presets.cavity(shape, lid_u=1.0)
    unf = unfinalized(shape[0], shape[1], g["kind"], g["bu"], g["bv"], delx=1 / 32, dely=1 / 32,
                      delt=2e-3, reynolds=100.0, sor_absolute_epsilon=eps, max_iterations=5000)
    rb = Simulation.try_from(unf, sor_mode=SOR_RED_BLACK, temporal_block=2)
    lex = Simulation.try_from(unf, sor_mode=SOR_REFERENCE_ORDER)
    for t in range(60):
        it_rb, n_rb = rb.run_simulation_tick()
        it_lex, n_lex = lex.run_simulation_tick()
        assert it_rb < 5000 and it_lex < 5000, (t, it_rb, it_lex, n_rb, n_lex)
    fluid = g["kind"] == 0
    du = np.abs(rb.grid.u - lex.grid.u)[fluid].max()
    dv = np.abs(rb.grid.v - lex.grid.v)[fluid].max()
    prb, plex = rb.grid.pressure, lex.grid.pressure
    dp = np.abs((prb - prb[fluid].mean()) - (plex - plex[fluid].mean()))[fluid].max()
    tol = eps
    assert np.abs(lex.grid.u)[fluid].max() > 0.05  # the lid has set the fluid in motion
    assert du <= tol and dv <= tol and dp <= tol, (du, dv, dp)


# ---- extensions: adaptive dt, moving wall -----------------------------------------------
def test_cavity_and_adaptive_dt_match_oracle():
    shape = (40, 40)
    g = presets.cavity(shape, lid_u=1.0)
    unf = unfinalized(shape[0], shape[1], g["kind"], g["bu"], g["bv"], delx=1 / 38,
                      dely=1 / 38, delt=1e-3, reynolds=1000.0, max_iterations=50)
    for mode, omode in ((SOR_REFERENCE_ORDER, po.SOR_REFERENCE_ORDER),
                        (SOR_RED_BLACK, po.SOR_RED_BLACK)):
        sim = Simulation.try_from(unf, sor_mode=mode, tau=0.5)
        o = oracle_from(unf, sor_mode=omode, tau=0.5)
        for t in range(8):
            it, nrm = sim.run_simulation_tick()
            oit, onrm = o.run_simulation_tick()
            assert it == oit and close(nrm, onrm), (t, it, oit, nrm, onrm)
            assert sim.delt == o.state().delt
        assert_bits_equal(sim.grid.u, o.u, "u")
        assert_bits_equal(sim.grid.v, o.v, "v")
        assert_bits_equal(sim.grid.pressure, o.p, "p")
        assert sim.time == o.state().time


# ---- device-side presets and cell edits --------------------------------------------------
def test_device_presets_match_host():
    cases = [("empty", (9, 7), ()), ("simple_inflow", (34, 18), ()), ("obstacle", (100, 20), ()),
             ("channel_circle", (120, 64), (40, 30, 9.5)), ("backward_step", (96, 48), (24, 20)),
             ("cavity", (32, 32), (1.5,))]
    for name, size, args in cases:
        if name == "channel_circle":
            host = presets.channel_circle(size, int(args[0]), int(args[1]), args[2])
        elif name == "backward_step":
            host = presets.backward_step(size, int(args[0]), int(args[1]))
        elif name == "cavity":
            host = presets.cavity(size, args[0])
        else:
            host = getattr(presets, name)(size)
        a = Simulation.from_preset(name, size, (0.1, 0.2), 0.005, 0.9, 100.0, 1e-3, 20, 1.7,
                                   preset_args=args)
        b = Simulation.try_from(unfinalized(size[0], size[1], host["kind"], host["bu"],
                                            host["bv"], max_iterations=20))
        assert np.array_equal(a.grid.cell_type, host["kind"]), name
        assert np.array_equal(a.grid.edge_type, b.grid.edge_type), name
        for _ in range(3):
            assert a.run_simulation_tick() == b.run_simulation_tick(), name
        assert_bits_equal(a.grid.u, b.grid.u, name)
        assert_bits_equal(a.grid.pressure, b.grid.pressure, name)


def test_draw_cells_and_rollback():
    """src/lib.rs:38-78: paint 2x2 blocks; a block that makes a wall too thin is rolled back
    and the previous boundary list stays in force."""
    shape = (40, 20)
    g = presets.simple_inflow(shape)
    unf = unfinalized(shape[0], shape[1], g["kind"], g["bu"], g["bv"])
    sim = Simulation.try_from(unf)
    sim.run_ticks(3)
    before = sim.grid.boundaries.sorted_boundary_list
    assert sim.grid.draw_cells(1, 10, 8) is True       # 2x2 NoSlip block: fine
    kind = g["kind"].copy()
    kind[10:12, 8:10] = 1
    assert np.array_equal(sim.grid.cell_type, kind)
    assert np.all(sim.grid.u[10:12, 8:10] == 0.0)
    after = sim.grid.boundaries.sorted_boundary_list
    assert len(after) == len(before) + 4
    # painting Fluid over half of it leaves a 1-wide wall -> rejected, rolled back
    edges = sim.grid.edge_type
    u_before = sim.grid.u
    assert sim.grid.draw_cells(0, 11, 7) is False
    assert np.array_equal(sim.grid.cell_type, kind)
    assert np.array_equal(sim.grid.edge_type, edges)
    assert_bits_equal(sim.grid.u, u_before, "u rolled back")
    assert sim.grid.boundaries.sorted_boundary_list == after
    # the oracle on the same edited mask agrees tick for tick afterwards
    o = oracle_from(unf)
    for _ in range(3):
        o.run_simulation_tick()
    o.kind[10:12, 8:10] = 1
    o.u[10:12, 8:10] = 0.0
    o.v[10:12, 8:10] = 0.0
    o.p[10:12, 8:10] = 0.0
    o.rebuild_boundary_list()
    for _ in range(3):
        assert sim.run_simulation_tick()[0] == o.run_simulation_tick()[0]
    assert_bits_equal(sim.grid.u, o.u, "u after edit")
    assert_bits_equal(sim.grid.pressure, o.p, "p after edit")


# ---- N4: colour mapping (src/visualization.rs) -- oracle-only parity (the reference has no
#      test for it), bit-exact RGBA8 --------------------------------------------------------
@pytest.mark.parametrize("shape,seed", [((34, 18), 51), ((100, 20), 52), ((257, 129), 53)])
def test_render_matches_oracle(shape, seed):
    nx, ny = shape
    kind, bu, bv = random_mask(nx, ny, seed)
    p, u, v = random_fields(nx, ny, seed)
    unf = unfinalized(nx, ny, kind, bu, bv, p=p, u=u, v=v)
    sim = Simulation.try_from(unf)
    o = oracle_from(unf)
    for tick in range(3):
        for ct in ("pressure", "speed"):
            a, b = sim.render_simulation(ct), o.render_simulation(ct)
            assert a.shape == (ny, nx, 4)
            assert np.array_equal(a, b), (ct, tick, np.argwhere(a != b)[:4])
        sim.run_simulation_tick()
        o.run_simulation_tick()
    # pressures outside a stale range (the range is only refreshed when SOR hits its cap):
    # hues below 0 and above 240 go through the same saturating casts
    sim.grid.pressure = p * 7.0
    o.p[:] = p * 7.0
    assert np.array_equal(sim.render_simulation("pressure"), o.render_simulation("pressure"))
    sim.close()


def test_render_degenerate_range():
    """fields at rest: range (0, 0) -> 0/0 = NaN hue -> `as u8` gives 0 (Rust saturating cast)"""
    nx, ny = 64, 32
    kind, bu, bv = po.preset("obstacle", nx, ny)
    unf = unfinalized(nx, ny, kind, bu, bv)
    sim = Simulation.try_from(unf)
    o = oracle_from(unf)
    for ct in ("pressure", "speed"):
        assert np.array_equal(sim.render_simulation(ct), o.render_simulation(ct))
    sim.close()


# ---- streaming kernel with boundary cells inside its items: straight walls along x and the
#      boundary rows at the ends of the grid (sor_rb_stream.cu, section 3a of DESIGN.md) ------
def _ring(nx, ny, kinds):
    """kind array with the given ring kinds (y-walls, x=0 row, x=nx-1 row); velocities 0 / inflow 1"""
    kind = np.zeros((nx, ny), dtype=np.uint8)
    bu = np.zeros((nx, ny))
    bv = np.zeros((nx, ny))
    wall, lo, hi = kinds
    kind[0, :] = lo
    kind[nx - 1, :] = hi
    kind[:, 0] = wall
    kind[:, ny - 1] = wall
    bu[kind == 3] = 1.0
    return kind, bu, bv


@pytest.mark.parametrize("T", [1, 2, 3, 4])
@pytest.mark.parametrize("shape,kinds,obstacle", [
    ((200, 300), (1, 1, 1), None),            # closed box: walls at both ends in x too
    ((260, 300), (1, 3, 2), None),            # channel
    ((260, 301), (1, 3, 2), None),            # odd NY: the last strip cannot start 16-byte aligned
    ((300, 120), (1, 3, 2), None),            # narrower than one strip
    ((240, 420), (1, 3, 2), (0, 40, 60, 90)),  # a block on the inflow row: not a plain end row
    ((240, 420), (1, 3, 2), (100, 0, 130, 50)),  # a block on the wall: the wall strip is cut
])
def test_red_black_walls_and_end_rows_in_stream(shape, kinds, obstacle, T):
    nx, ny = shape
    kind, bu, bv = _ring(nx, ny, kinds)
    if obstacle:
        x0, y0, x1, y1 = obstacle
        kind[x0:x1, y0:y1] = 1
        bu[x0:x1, y0:y1] = 0.0
    p, u, v = random_fields(nx, ny, 77 + T)
    unf = unfinalized(nx, ny, kind, bu, bv, p=p, u=u, v=v)
    sim = Simulation.try_from(unf, sor_mode=SOR_RED_BLACK, temporal_block=T)
    o = oracle_from(unf, sor_mode=po.SOR_RED_BLACK)
    n = 2 * T + 3   # full passes and a shortened last one
    norms = sim.sor_sweeps(n)
    for k in range(n):
        o.sor_sweep()
        assert close(norms[k], o.calculate_norm_squared()), (k, norms[k])
    assert_bits_equal(sim.grid.pressure, o.p, "p after red-black sweeps")
    for t in range(2):
        it, nrm = sim.run_simulation_tick()
        oit, onrm = o.run_simulation_tick()
        assert it == oit and close(nrm, onrm), (t, it, oit, nrm, onrm)
    assert_bits_equal(sim.grid.pressure, o.p, "p after ticks")
    assert_bits_equal(sim.grid.u, o.u, "u after ticks")
    assert sim.grid.pressure_range == list(o.state().pressure_range)
    assert sim.grid.speed_range == list(o.state().speed_range)
    sim.close()


def test_stream_kinds_env_gives_identical_fields(monkeypatch):
    """whatever subset of item kinds the streaming kernel takes, the fields are the same bits"""
    monkeypatch.setenv("SB_SOR_MID", "0")   # this is about the pass kernels
    nx, ny = 300, 380
    kind, bu, bv = _ring(nx, ny, (1, 3, 2))
    p, u, v = random_fields(nx, ny, 5)
    unf = unfinalized(nx, ny, kind, bu, bv, p=p, u=u, v=v)
    ref = None
    for kinds in ("0", "1", "2", "3"):
        monkeypatch.setenv("SB_RB_STREAM_KINDS", kinds)
        sim = Simulation.try_from(unf, sor_mode=SOR_RED_BLACK, temporal_block=4)
        for _ in range(2):
            sim.run_simulation_tick()
        got = (sim.grid.pressure, sim.grid.u, sim.grid.v, sim.rb_plan)
        sim.close()
        if ref is None:
            ref = got
            assert got[3][1] > 0          # some streaming items even with plain strips only
        else:
            for a, b, name in zip(ref[:3], got[:3], "puv"):
                assert_bits_equal(a, b, f"{name} with kinds {kinds}")
    assert got[3][0] < ref[3][0]          # all kinds: fewer tiles left for the tile kernel


# ---- the grid-resident cooperative solve (sor_mid.cu) on forced decompositions ------------
@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("ctas", [1, 2, 3, 7, 64, 1000])
@pytest.mark.parametrize("shape,blocks,seed", [((257, 129), 6, 61), ((150, 300), 3, 62),
                                               ((9, 40), 0, 63), ((331, 64), 4, 64),
                                               ((64, 1023), 2, 65)])
def test_mid_kernel_decompositions_match_oracle(shape, blocks, seed, ctas, variant, monkeypatch):
    """One cooperative launch per solve, the grid split into `ctas` bands of x-rows (down to
    2 rows per band; if a band would not fit an SM, more bands) that exchange edge rows through
    L2 -- variant 1: generic kernel, two rows per half-sweep; variant 2: register-window
    kernel, four rows per sweep: p bit-exact against the
    oracle's red-black restatement whatever the decomposition, norms of every sweep within
    the summation-order allowance, obstacles straddling band edges included; then full ticks
    (exit rule decided on the device by every CTA alike)."""
    if SOR_PATH != "mid-kernel":
        pytest.skip("decompositions are forced once")
    monkeypatch.setenv("SB_SOR_MID_CTAS", str(ctas))
    monkeypatch.setenv("SB_SOR_MID_VARIANT", str(variant))
    nx, ny = shape
    kind, bu, bv = random_mask(nx, ny, seed, n_blocks=blocks)
    p, u, v = random_fields(nx, ny, seed)
    unf = unfinalized(nx, ny, kind, bu, bv, p=p, u=u, v=v)
    sim = Simulation.try_from(unf, sor_mode=SOR_RED_BLACK)
    o = oracle_from(unf, sor_mode=po.SOR_RED_BLACK)
    n = 9
    launches = sim.kernel_launches
    norms = sim.sor_sweeps(n)
    assert sim.kernel_launches - launches == 1, "the whole solve is one launch"
    assert sim.sor_path[0] == 1 + variant, sim.sor_path
    for k in range(n):
        o.sor_sweep()
        assert close(norms[k], o.calculate_norm_squared()), (k, norms[k])
    assert_bits_equal(sim.grid.pressure, o.p, "p after red-black sweeps")
    for t in range(3):
        it, nrm = sim.run_simulation_tick()
        oit, onrm = o.run_simulation_tick()
        assert it == oit and close(nrm, onrm), (t, it, oit, nrm, onrm)
    assert_bits_equal(sim.grid.pressure, o.p, "p after ticks")
    assert_bits_equal(sim.grid.u, o.u, "u after ticks")
    assert_bits_equal(sim.grid.v, o.v, "v after ticks")


def test_mid_kernel_early_exit():
    """A grid of mid size whose solves end by the exit rule after 1 .. 100 sweeps (coarse
    eps): all CTAs must leave the loop in the same sweep, and the pressure range is only
    recomputed after a capped solve (simulation.rs:283)."""
    if SOR_PATH == "pass-kernels":
        pytest.skip("covers the resident kernels")
    shape = (140, 96)
    g = presets.simple_inflow(shape)
    unf = unfinalized(shape[0], shape[1], g["kind"], g["bu"], g["bv"], sor_absolute_epsilon=50.0)
    sim = Simulation.try_from(unf, sor_mode=SOR_RED_BLACK)
    o = oracle_from(unf, sor_mode=po.SOR_RED_BLACK)
    seen = set()
    for t in range(30):
        it, nrm = sim.run_simulation_tick()
        oit, onrm = o.run_simulation_tick()
        assert it == oit and close(nrm, onrm), (t, it, oit, nrm, onrm)
        seen.add(it)
    assert len(seen) > 1, seen
    assert_bits_equal(sim.grid.pressure, o.p, "p")
    assert_bits_equal(sim.grid.u, o.u, "u")
    assert list(sim.grid.pressure_range) == list(o.state().pressure_range)


# ---- frozen tiles: deep inside an obstacle no pass touches anything ------------------------
@pytest.mark.parametrize("T", [1, 2, 3, 4])
def test_frozen_tiles_inside_a_block_match_oracle(T, monkeypatch):
    """A large NoSlip block (the backward-facing step of BASELINE config 4 in small): the tiles
    deep inside it are left out of every pass -- their pressures are mirrored into the other
    buffer once per solve and their residual sum (the reference's norm counts obstacle
    cells, simulation.rs:216-227) is added as a constant.  Fields bit-exact vs the oracle,
    norms within the summation-order allowance, and identical to the run that keeps those
    tiles on the tile kernel."""
    if SOR_PATH != "pass-kernels":
        pytest.skip("about the pass kernels")
    nx, ny = 300, 520
    kind, bu, bv = po.preset("simple_inflow", nx, ny)
    kind = kind.copy()
    kind[40:260, 130:500] = 1          # NoSlip block, away from the ring
    p, u, v = random_fields(nx, ny, 71)   # random p inside the block too: residuals there != 0
    unf = unfinalized(nx, ny, kind, bu, bv, p=p, u=u, v=v)
    o = oracle_from(unf, sor_mode=po.SOR_RED_BLACK)
    got = {}
    for frozen in ("1", "0"):
        monkeypatch.setenv("SB_RB_FROZEN", frozen)
        sim = Simulation.try_from(unf, sor_mode=SOR_RED_BLACK, temporal_block=T)
        norms = sim.sor_sweeps(7)
        plan = sim.rb_plan
        ticks = [sim.run_simulation_tick() for _ in range(2)]
        got[frozen] = (norms, plan, ticks, sim.grid.pressure, sim.grid.u, sim.grid.v)
        sim.close()
    assert got["1"][1][0] < got["0"][1][0], (got["1"][1], got["0"][1])   # fewer tile-kernel tiles
    norms = got["1"][0]
    for k in range(7):
        o.sor_sweep()
        assert close(norms[k], o.calculate_norm_squared()), (k, norms[k])
        assert close(norms[k], got["0"][0][k])
    for t in range(2):
        oit, onrm = o.run_simulation_tick()
        it, nrm = got["1"][2][t]
        assert it == oit and close(nrm, onrm), (t, it, oit, nrm, onrm)
    for i, name in ((3, "p"), (4, "u"), (5, "v")):
        assert_bits_equal(got["1"][i], got["0"][i], name + " frozen vs tile kernel")
    assert_bits_equal(got["1"][3], o.p, "p")
    assert_bits_equal(got["1"][4], o.u, "u")


# ---- behaviours of the mirrored pub surface outside the tick order ---------------------------
@pytest.mark.parametrize("mode,omode", [(SOR_REFERENCE_ORDER, po.SOR_REFERENCE_ORDER),
                                        (SOR_RED_BLACK, po.SOR_RED_BLACK)])
def test_set_u_and_v_before_any_velocity_bc(mode, omode):
    """pub fn set_u_and_v straight after construction / rebuild_boundary_list: u_v_restore is
    empty (src/grid/mod.rs:205), so nothing is restored over the boundary cells."""
    nx, ny = 64, 48
    kind, bu, bv = random_mask(nx, ny, 81)
    p, u, v = random_fields(nx, ny, 81)
    unf = unfinalized(nx, ny, kind, bu, bv, p=p, u=u, v=v)
    sim = Simulation.try_from(unf, sor_mode=mode)
    o = oracle_from(unf, sor_mode=omode)
    sim.set_u_and_v()
    o.set_u_and_v()
    assert_bits_equal(sim.grid.u, o.u, "u: set_u_and_v after try_from")
    assert_bits_equal(sim.grid.v, o.v, "v: set_u_and_v after try_from")
    # a tick fills the restore records; a rebuild empties them again
    sim.run_simulation_tick()
    o.run_simulation_tick()
    sim.grid.rebuild_boundary_list()
    o.rebuild_boundary_list()
    sim.set_u_and_v()
    o.set_u_and_v()
    assert_bits_equal(sim.grid.u, o.u, "u: set_u_and_v after rebuild")
    assert_bits_equal(sim.grid.v, o.v, "v: set_u_and_v after rebuild")
    assert sim.grid.speed_range == list(o.state().speed_range)
    sim.close()


@pytest.mark.parametrize("mode,omode", [(SOR_REFERENCE_ORDER, po.SOR_REFERENCE_ORDER),
                                        (SOR_RED_BLACK, po.SOR_RED_BLACK)])
def test_initial_norm_none_is_latched_after_the_first_sweep(mode, omode):
    """sim.initial_norm_squared = None: solve_sor caches the norm after its first sweep
    (src/simulation.rs:229-237, 276), so iteration 0 can only leave through the eps test."""
    shape = (40, 24)
    g = presets.simple_inflow(shape)
    p, u, v = random_fields(shape[0], shape[1], 83, scale=0.3)
    unf = unfinalized(shape[0], shape[1], g["kind"], g["bu"], g["bv"], p=p, u=u, v=v,
                      max_iterations=30)
    sim = Simulation.try_from(unf, sor_mode=mode)
    o = oracle_from(unf, sor_mode=omode)
    for t in range(4):
        if t in (0, 2):
            sim.initial_norm_squared = None
            o.clear_initial_norm()
            assert sim.initial_norm_squared is None
        it, nrm = sim.run_simulation_tick()
        oit, onrm = o.run_simulation_tick()
        assert it == oit and close(nrm, onrm), (t, it, oit, nrm, onrm)
        assert close(sim.initial_norm_squared, o.state().initial_norm_squared)
        assert_bits_equal(sim.grid.pressure, o.p, f"p tick {t}")
        assert_bits_equal(sim.grid.u, o.u, f"u tick {t}")
    sim.close()


def test_boundary_velocity_table_replaces_the_old_one():
    """sb_set_boundary_velocities with a shrunken / empty table: cells it no longer names
    fall back to velocity (0, 0) instead of keeping their previous inflow velocity."""
    import ctypes as C

    from stroemung_b200 import _capi
    shape = (40, 20)
    g = presets.simple_inflow(shape)
    unf = unfinalized(shape[0], shape[1], g["kind"], g["bu"], g["bv"])
    sim = Simulation.try_from(unf)
    sim.run_ticks(2)
    keep = [(0, y, 0.5, 0.0) for y in range(1, 10)]       # the other inflow cells: dropped
    tab = (_capi.BoundaryVelocity * len(keep))(*[_capi.BoundaryVelocity(*e) for e in keep])
    sim._check(_capi.lib().sb_set_boundary_velocities(sim._h, tab, len(keep)))
    bu = np.zeros(shape)
    for x, y, uu, _ in keep:
        bu[x, y] = uu
    o = oracle_from(unf)
    for _ in range(2):
        o.run_simulation_tick()
    o.bu[:] = bu
    for _ in range(2):
        assert sim.run_simulation_tick()[0] == o.run_simulation_tick()[0]
    assert_bits_equal(sim.grid.u, o.u, "u after a shrunken table")
    sim._check(_capi.lib().sb_set_boundary_velocities(sim._h, None, 0))
    o.bu[:] = 0.0
    for _ in range(2):
        assert sim.run_simulation_tick()[0] == o.run_simulation_tick()[0]
    assert_bits_equal(sim.grid.u, o.u, "u after an empty table")
    assert_bits_equal(sim.grid.pressure, o.p, "p after an empty table")
    sim.close()


def test_adaptive_dt_after_field_upload():
    """tau > 0: delt comes from the maxima of the CURRENT u, v (the oracle recomputes them at
    the start of the tick), also when the host has replaced the fields since the last tick."""
    shape = (40, 40)
    g = presets.cavity(shape, lid_u=1.0)
    unf = unfinalized(shape[0], shape[1], g["kind"], g["bu"], g["bv"], delx=1 / 38, dely=1 / 38,
                      delt=1e-3, reynolds=1000.0, max_iterations=30)
    sim = Simulation.try_from(unf, sor_mode=SOR_RED_BLACK, tau=0.5)
    o = oracle_from(unf, sor_mode=po.SOR_RED_BLACK, tau=0.5)
    for _ in range(3):
        sim.run_simulation_tick()
        o.run_simulation_tick()
    _, u, v = random_fields(shape[0], shape[1], 85, scale=3.0)
    sim.grid.u = u
    sim.grid.v = v
    o.u[:] = u
    o.v[:] = v
    for t in range(2):
        it, nrm = sim.run_simulation_tick()
        oit, onrm = o.run_simulation_tick()
        assert sim.delt == o.state().delt, (t, sim.delt, o.state().delt)
        assert it == oit and close(nrm, onrm)
    assert_bits_equal(sim.grid.u, o.u, "u")
    sim.close()
