"""The C oracle cross-checked against a second, independent restatement of the reference
(tests/py_restatement.py: pure Python, written from the Rust sources) -- first the Python
restatement itself is pinned on the reference's snapshots, then both run cases the reference's
own tests never reach (grids beyond 5x7, obstacles, every corner edge class,
interior Inflow / Outflow cells) and must agree bit for bit."""
import numpy as np
import pytest

from oracle import pyoracle as po
from stroemung_b200 import refjson
from tests import py_restatement as pr
from tests.util import DEFAULTS, assert_bits_equal, random_fields, random_mask

TICK = "stroemung__simulation__tests__simulation_tick"
PRM = dict(delx=0.1, dely=0.2, delt=0.005, gamma=0.9, reynolds=100.0, sor_absolute_epsilon=0.001,
           max_iterations=100, omega=1.7)


def arr(a):
    return np.array(a, dtype=np.float64)


def test_python_restatement_reproduces_the_tick_snapshots(kat, snapshots):
    """src/simulation.rs:571-618 with its nine snapshots, through the Python restatement"""
    kind, bu, bv = po.preset("simple_inflow", 4, 3)
    z = np.zeros((4, 3))
    sim = pr.PySim(4, 3, kind, bu, bv, z, z, z, **PRM)
    it, nrm = sim.run_simulation_tick()
    for suffix, field in (("", sim.f), ("-2", sim.g), ("-3", sim.rhs)):
        assert_bits_equal(arr(field), refjson.array_from_json(snapshots[TICK + suffix]["json"]),
                          "tick1" + suffix)
    a = kat["simulation_tick_asserts"]
    assert (it, nrm) == (a[0]["sor_iterations"], a[0]["norm_squared"])

    def check(snap):
        assert_bits_equal(arr(sim.p), refjson.array_from_json(snap["grid"]["pressure"]), "p")
        assert_bits_equal(arr(sim.u), refjson.array_from_json(snap["grid"]["u"]), "u")
        assert_bits_equal(arr(sim.v), refjson.array_from_json(snap["grid"]["v"]), "v")
        assert sim.time == snap["time"] and sim.iterations == snap["iterations"]
        assert sim.initial_norm_squared == snap["initial_norm_squared"]
    check(snapshots[TICK + "-4"]["json"])
    for _ in range(100):
        it, nrm = sim.run_simulation_tick()
    assert (it, nrm) == (a[1]["sor_iterations"], a[1]["norm_squared"])
    for suffix, field in (("-5", sim.f), ("-6", sim.g), ("-7", sim.rhs)):
        assert_bits_equal(arr(field), refjson.array_from_json(snapshots[TICK + suffix]["json"]),
                          "tick101" + suffix)
    check(snapshots[TICK + "-8"]["json"])
    for _ in range(100):
        sim.run_simulation_tick()
    check(snapshots[TICK + "-9"]["json"])


@pytest.mark.parametrize("shape,seed,blocks", [((14, 11), 1, 2), ((24, 18), 2, 5), ((31, 16), 3, 6),
                                               ((20, 27), 4, 6), ((26, 22), 5, 8)])
def test_c_oracle_equals_python_restatement_on_random_grids(shape, seed, blocks):
    nx, ny = shape
    kind, bu, bv = random_mask(nx, ny, seed, n_blocks=blocks)
    p, u, v = random_fields(nx, ny, seed)
    prm = dict(DEFAULTS, max_iterations=12)
    o = po.OracleSim(nx, ny, kind=kind, p=p, u=u, v=v, bu=bu, bv=bv, **prm)
    s = pr.PySim(nx, ny, kind, bu, bv, p, u, v, **prm)
    # classification: same list, same order, same edge classes, same fluid count
    idx, edge = o.boundary_list()
    assert [(int(i) // ny, int(i) % ny) for i in idx] == [c for c, _ in s.blist]
    assert [int(e) for e in edge] == [e for _, e in s.blist]
    edges_seen = {e for _, e in s.blist}
    st = o.state()
    assert st.fluid_cells == s.fluid_cells
    assert st.initial_norm_squared == s.initial_norm_squared
    assert list(st.pressure_range) == s.pressure_range and list(st.speed_range) == s.speed_range
    for name in ("f", "g", "rhs"):
        assert_bits_equal(getattr(o, name), arr(getattr(s, name)), name + " at construction")
    for t in range(4):
        # stage by stage on the first tick, whole ticks afterwards
        if t == 0:
            o.set_boundary_u_and_v(); s.set_boundary_u_and_v()
            assert_bits_equal(o.u, arr(s.u), "u after BC"); assert_bits_equal(o.v, arr(s.v), "v after BC")
            o.calculate_f_and_g(); s.calculate_f_and_g()
            assert_bits_equal(o.f, arr(s.f), "f"); assert_bits_equal(o.g, arr(s.g), "g")
            o.calculate_rhs(); s.calculate_rhs()
            assert_bits_equal(o.rhs, arr(s.rhs), "rhs")
            got, want = o.solve_sor(), s.solve_sor()
            assert got == want, (got, want)
            assert_bits_equal(o.p, arr(s.p), "p after SOR")
            o.set_u_and_v(); s.set_u_and_v()
        else:
            got, want = o.run_simulation_tick(), s.run_simulation_tick()
            assert got == want, (t, got, want)
        for name in ("p", "u", "v", "f", "g", "rhs"):
            assert_bits_equal(getattr(o, name), arr(getattr(s, name)), f"{name} after tick {t}")
        st = o.state()
        assert list(st.speed_range) == s.speed_range
        assert list(st.pressure_range) == s.pressure_range
    assert len(edges_seen) >= 7, edges_seen   # corners and flat edges all occur


def test_thin_boundary_is_the_first_in_x_major_order():
    nx, ny = 12, 10
    kind, bu, bv = po.preset("simple_inflow", nx, ny)
    kind = kind.copy()
    kind[7, 3:7] = 1      # one cell wide along y: fluid on both x sides
    kind[4, 5] = 1        # a lone cell further up in x-major order
    with pytest.raises(po.BoundaryTooThin) as e:
        po.OracleSim(nx, ny, kind=kind, bu=bu, bv=bv, **DEFAULTS)
    with pytest.raises(pr.TooThin) as e2:
        pr.classify(kind.tolist(), nx, ny)
    assert tuple(int(c) for c in e.value.xy) == tuple(e2.value.args[0]) == (4, 5)


def test_obstacle_preset_default_config_three_ticks():
    """BASELINE config 1 (presets::obstacle([100, 20]), CLI defaults of src/args.rs): the residual
    norm is dominated by the obstacle's cells, so every solve runs max_iterations = 100 sweeps
     -- in both restatements, with identical bits."""
    nx, ny = 100, 20
    kind, bu, bv = po.preset("obstacle", nx, ny)
    z = np.zeros((nx, ny))
    o = po.OracleSim(nx, ny, kind=kind, bu=bu, bv=bv, **PRM)
    s = pr.PySim(nx, ny, kind, bu, bv, z, z, z, **PRM)
    assert o.state().initial_norm_squared == s.initial_norm_squared == 0.0
    for t in range(3):
        got, want = o.run_simulation_tick(), s.run_simulation_tick()
        assert got == want and got[0] == 100, (t, got, want)
    for name in ("p", "u", "v", "f", "g", "rhs"):
        assert_bits_equal(getattr(o, name), arr(getattr(s, name)), name)
    st = o.state()
    assert list(st.pressure_range) == s.pressure_range and list(st.speed_range) == s.speed_range


@pytest.mark.parametrize("eps", [1e-3, 1e-6])
def test_oracle_red_black_agrees_with_reference_order_when_converged(eps):
    """Performance-mode parity bar at the oracle level (BASELINE north star): on
    converged ticks the red-black ordering gives the reference-order fields within the SOR
    tolerance -- u, v and the mean-removed p, fluid cells only.  A closed lid-driven cavity has a
    solvable pressure problem every tick (the channel presets hit the sweep cap instead)."""
    from stroemung_b200 import presets
    from tests.util import oracle_from, unfinalized
    shape = (34, 34)
    g = presets.cavity(shape, lid_u=1.0)
    unf = unfinalized(shape[0], shape[1], g["kind"], g["bu"], g["bv"], delx=1 / 32, dely=1 / 32,
                      delt=2e-3, reynolds=100.0, sor_absolute_epsilon=eps, max_iterations=5000)
    rb = oracle_from(unf, sor_mode=po.SOR_RED_BLACK)
    lex = oracle_from(unf, sor_mode=po.SOR_REFERENCE_ORDER)
    for t in range(60):
        it_rb, n_rb = rb.run_simulation_tick()
        it_lex, n_lex = lex.run_simulation_tick()
        assert it_rb < 5000 and it_lex < 5000, (t, it_rb, it_lex, n_rb, n_lex)
    fluid = g["kind"] == 0
    du = np.abs(rb.u - lex.u)[fluid].max()
    dv = np.abs(rb.v - lex.v)[fluid].max()
    dp = np.abs((rb.p - rb.p[fluid].mean()) - (lex.p - lex.p[fluid].mean()))[fluid].max()
    assert np.abs(lex.u)[fluid].max() > 0.05   # the lid has set the fluid in motion
    tol = 3.0 * eps
    assert du <= tol and dv <= tol and dp <= tol, (du, dv, dp)
