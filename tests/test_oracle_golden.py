"""Pin the CPU oracle on the reference's own golden vectors.

Everything here is exact f64 equality (bit patterns, so -0.0 != 0.0), as in the
reference's `assert_eq!` / insta snapshot tests.
"""
import numpy as np
import pytest

from oracle import pyoracle as po
from stroemung_b200 import nast2d, refjson

TICK = "stroemung__simulation__tests__simulation_tick"


def bits(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64)).view(np.uint64)


def assert_bits_equal(a, b, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    bad = np.nonzero(bits(a) != bits(b))
    assert bad[0].size == 0, f"{what}: {bad[0].size} mismatches, first at " \
        f"{tuple(x[0] for x in bad)}: {a[tuple(x[0] for x in bad)]!r} vs " \
        f"{b[tuple(x[0] for x in bad)]!r}"


# ---- 32 KATs: src/math.rs:192-402, src/simulation.rs:449-569 -------------------
def test_kat_du2dx(kat):
    for c in kat["du2dx"]:
        assert po.du2dx(c["u"], c["delx"], c["gamma"]) == c["expected"]


def test_kat_dv2dy(kat):
    for c in kat["dv2dy"]:
        assert po.dv2dy(c["v"], c["dely"], c["gamma"]) == c["expected"]


def test_kat_duvdx(kat):
    for c in kat["duvdx"]:
        assert po.duvdx(c["u"], c["v"], c["delx"], c["gamma"]) == c["expected"]


def test_kat_duvdy(kat):
    for c in kat["duvdy"]:
        assert po.duvdy(c["u"], c["v"], c["dely"], c["gamma"]) == c["expected"]


def test_kat_laplacian(kat):
    for c in kat["laplacian"]:
        assert po.laplacian(c["e"], c["delx"], c["dely"]) == c["expected"]


def test_kat_calculate_f_g(kat):
    for name, fn in (("calculate_f", po.calculate_f), ("calculate_g", po.calculate_g)):
        for c in kat[name]:
            got = fn(c["u"], c["v"], c["delx"], c["dely"], c["delt"], c["gamma"],
                     c["reynolds"])
            assert got == c["expected"], (name, got, c["expected"])


# ---- simulation_tick: src/simulation.rs:571-618 + nine snapshots -------------
def make_tick_sim(sor_mode=po.SOR_REFERENCE_ORDER):
    kind, bu, bv = po.preset("simple_inflow", 4, 3)
    return po.OracleSim(4, 3, delx=0.1, dely=0.2, delt=0.005, gamma=0.9, reynolds=100.0,
                        sor_absolute_epsilon=0.001, max_iterations=100, omega=1.7,
                        kind=kind, bu=bu, bv=bv, sor_mode=sor_mode)


def check_sim_snapshot(sim, snap):
    assert_bits_equal(sim.p, refjson.array_from_json(snap["grid"]["pressure"]), "pressure")
    assert_bits_equal(sim.u, refjson.array_from_json(snap["grid"]["u"]), "u")
    assert_bits_equal(sim.v, refjson.array_from_json(snap["grid"]["v"]), "v")
    st = sim.state()
    assert st.time == snap["time"]
    assert st.iterations == snap["iterations"]
    assert st.has_initial_norm == 1
    assert st.initial_norm_squared == snap["initial_norm_squared"]
    kind, bu, bv = refjson.cells_from_json(snap["grid"]["cell_type"])
    assert np.array_equal(sim.kind, kind)


def test_simulation_tick_snapshots(kat, snapshots):
    sim = make_tick_sim()
    it, nrm = sim.run_simulation_tick()
    for suffix, field in (("", sim.f), ("-2", sim.g), ("-3", sim.rhs)):
        assert_bits_equal(field, refjson.array_from_json(snapshots[TICK + suffix]["json"]),
                          "tick1" + suffix)
    check_sim_snapshot(sim, snapshots[TICK + "-4"]["json"])
    a = kat["simulation_tick_asserts"]
    assert (it, nrm) == (a[0]["sor_iterations"], a[0]["norm_squared"])
    for _ in range(100):
        it, nrm = sim.run_simulation_tick()
    assert (it, nrm) == (a[1]["sor_iterations"], a[1]["norm_squared"])
    for suffix, field in (("-5", sim.f), ("-6", sim.g), ("-7", sim.rhs)):
        assert_bits_equal(field, refjson.array_from_json(snapshots[TICK + suffix]["json"]),
                          "tick101" + suffix)
    check_sim_snapshot(sim, snapshots[TICK + "-8"]["json"])
    for _ in range(100):
        sim.run_simulation_tick()
    check_sim_snapshot(sim, snapshots[TICK + "-9"]["json"])


# ---- construction: src/simulation.rs:408-447 ----------------------------------
def sim_from_docs(prm, grid, **kw):
    return po.OracleSim(prm["size"][0], prm["size"][1], delx=prm["cell_size"][0],
                        dely=prm["cell_size"][1], delt=prm["delt"], gamma=prm["gamma"],
                        reynolds=prm["reynolds"],
                        sor_absolute_epsilon=prm["sor_absolute_epsilon"],
                        max_iterations=prm["max_iterations"], omega=prm["omega"],
                        kind=grid["kind"], p=grid["p"], u=grid["u"], v=grid["v"],
                        bu=grid["bu"], bv=grid["bv"],
                        initial_norm_squared=prm["initial_norm_squared"],
                        iterations=prm["iterations"], time=prm["time"], **kw)


def test_deserialize_initial_norm(fixtures, snapshots):
    # 5x7 all-fluid, all-zero: initial_norm_squared == 0.0
    prm, grid = refjson.simulation_from_json(
        fixtures["src/test_data/simple_simulation.json"]["json"])
    sim = sim_from_docs(prm, grid)
    snap = snapshots["stroemung__simulation__tests__deserialize"]["json"]
    assert sim.state().initial_norm_squared == snap["initial_norm_squared"] == 0.0
    # NaSt2D-derived 4x3 state: 899.9547140394143, with the reference parser's 1-ulp quirk
    raw = fixtures["src/test_data/small_simulation_with_boundaries.json"]["raw"]
    prm, grid = refjson.simulation_from_json(refjson.loads(raw, quirk_serde_json=True))
    snap = snapshots["stroemung__simulation__tests__deserialize-2"]["json"]
    assert_bits_equal(grid["p"], refjson.array_from_json(snap["grid"]["pressure"]), "parsed p")
    sim = sim_from_docs(prm, grid)
    assert snap["initial_norm_squared"] == 899.9547140394143
    assert sim.state().initial_norm_squared == snap["initial_norm_squared"]
    check_sim_snapshot(sim, snap)
    # a correctly-rounded parse gives the neighbouring value (documents the quirk)
    prm2, grid2 = refjson.simulation_from_json(refjson.loads(raw))
    assert sim_from_docs(prm2, grid2).state().initial_norm_squared == 899.9547140394145


def test_serialize_snapshot(snapshots):
    # src/simulation.rs:422-447: presets::empty([5,7]) -> all zero, norm 0
    kind, bu, bv = po.preset("empty", 5, 7)
    sim = po.OracleSim(5, 7, delx=1.0, dely=2.0, delt=1.4, gamma=1.7, reynolds=100.0,
                       sor_absolute_epsilon=0.001, max_iterations=100, omega=1.7, kind=kind)
    check_sim_snapshot(sim, snapshots["stroemung__simulation__tests__serialize"]["json"])
    doc = refjson.simulation_to_json(
        {"size": (5, 7), "cell_size": (1.0, 2.0), "delt": 1.4, "gamma": 1.7, "reynolds": 100.0,
         "initial_norm_squared": sim.state().initial_norm_squared,
         "sor_absolute_epsilon": 0.001, "max_iterations": 100, "iterations": 0, "time": 0.0,
         "omega": 1.7},
        {"p": sim.p, "u": sim.u, "v": sim.v, "kind": kind, "bu": bu, "bv": bv})
    assert doc == snapshots["stroemung__simulation__tests__serialize"]["json"]


# ---- boundary classification (integer): src/grid/mod.rs:683-823 ---------------
def grid3(cells):
    kind = np.zeros((3, 3), dtype=np.uint8)
    for c in cells:
        kind[c] = po.KIND_NOSLIP
    return kind


def make3(kind):
    return po.OracleSim(3, 3, delx=1.0, dely=1.0, delt=0.1, gamma=0.9, reynolds=100.0,
                        sor_absolute_epsilon=1e-3, max_iterations=1, omega=1.7, kind=kind)


def test_thin_boundary():
    # src/grid/mod.rs:683-705
    for cells in ([(1, 0), (1, 1), (1, 2)], [(0, 1), (1, 1), (2, 1)]):
        with pytest.raises(po.BoundaryTooThin):
            make3(grid3(cells))


def test_rebuild_boundary_list():
    # src/grid/mod.rs:707-801
    N, NE, E, SE, S, SW, W, NW = range(1, 9)
    examples = [
        ([(0, 0), (0, 1), (0, 2), (1, 0), (1, 2), (2, 0), (2, 1), (2, 2)],
         [0, E, 0, S, N, 0, W, 0]),
        ([(0, 0), (0, 2), (2, 0), (2, 2)], [SE, NE, SW, NW]),
    ]
    for cells, edges in examples:
        sim = make3(grid3(cells))
        idx, edge = sim.boundary_list()
        assert [(int(i) // 3, int(i) % 3) for i in idx] == cells
        assert list(edge) == edges
        assert sim.state().fluid_cells == 9 - len(cells)


def test_deserialize_boundaries_snapshot(fixtures, snapshots):
    # src/grid/mod.rs:813-823 + deserialize_boundaries-2.snap (Display of BoundaryList)
    grid = refjson.grid_from_json(fixtures["src/test_data/small_grid_with_boundaries.json"]["json"])
    sim = po.OracleSim(4, 3, delx=0.1, dely=0.2, delt=0.005, gamma=0.9, reynolds=100.0,
                       sor_absolute_epsilon=1e-3, max_iterations=100, omega=1.7,
                       kind=grid["kind"], p=grid["p"], u=grid["u"], v=grid["v"],
                       bu=grid["bu"], bv=grid["bv"])
    idx, edge = sim.boundary_list()
    lines = ["Boundaries:"]
    lines += [f"  BoundaryIndex({int(i) // 3}, {int(i) % 3})" for i in idx]
    lines.append("Sorted Boundary List:")
    nb = {"North": lambda x, y: [("north", (x, y - 1))],
          "NorthEast": lambda x, y: [("north", (x, y - 1)), ("east", (x + 1, y))],
          "East": lambda x, y: [("east", (x + 1, y))],
          "SouthEast": lambda x, y: [("south", (x, y + 1)), ("east", (x + 1, y))],
          "South": lambda x, y: [("south", (x, y + 1))],
          "SouthWest": lambda x, y: [("south", (x, y + 1)), ("west", (x - 1, y))],
          "West": lambda x, y: [("west", (x - 1, y))],
          "NorthWest": lambda x, y: [("north", (x, y - 1)), ("west", (x - 1, y))]}
    for i, e in zip(idx, edge):
        x, y = int(i) // 3, int(i) % 3
        if e == 0:
            lines.append(f"  (({x}, {y}), None)")
        else:
            name = po.EDGE_NAMES[e]
            fields = ", ".join(f"{d}_neighbor: ({a}, {b})" for d, (a, b) in nb[name](x, y))
            lines.append(f"  (({x}, {y}), Some({name} {{ {fields} }}))")
    want = snapshots["stroemung__grid__tests__deserialize_boundaries-2"]["text"]
    assert "\n".join(lines).strip() == want.strip()


# ---- NaSt2D test data re-validated: python/test_generate_test_data.py:28-49 ------
def test_nast2d_fixture(fixtures, snapshots):
    raw = bytes.fromhex(fixtures["python/test_data/small_data.out"]["hex"])
    assert len(raw) == 440
    out = nast2d.parse_out(raw)
    exp = fixtures["python/test_data/small_data.out_expected.json"]["json"]
    assert (out["imax"], out["jmax"]) == (exp["imax"], exp["jmax"]) == (2, 1)
    for k in ("U", "V", "P", "T"):
        assert_bits_equal(out[k], np.array(exp[k]), k)
    assert np.array_equal(out["flags"] & 16, np.array(exp["flags"]) & 16)
    grid = nast2d.as_grid(out)
    rust = refjson.grid_from_json(
        fixtures["python/test_data/small_data.out_rust_expected.json"]["json"])
    for k in ("p", "u", "v", "bu", "bv"):
        assert_bits_equal(grid[k], rust[k], k)
    assert np.array_equal(grid["kind"], rust["kind"])
    # tests/test_file_parsing.rs:15-21: the converted grid classifies without error
    assert fixtures["tests/test_data/small_data.out.json"]["json"] == \
        fixtures["python/test_data/small_data.out_rust_expected.json"]["json"]
    sim = po.OracleSim(4, 3, delx=0.1, dely=0.2, delt=0.005, gamma=0.9, reynolds=100.0,
                       sor_absolute_epsilon=1e-3, max_iterations=100, omega=1.7,
                       kind=grid["kind"], p=grid["p"], u=grid["u"], v=grid["v"],
                       bu=grid["bu"], bv=grid["bv"])
    snap = snapshots["test_file_parsing__deserialize"]["json"]
    # (the snapshot holds the reference parser's 1-ulp-low pressure literal)
    assert np.array_equal(sim.kind, refjson.cells_from_json(snap["cell_type"])[0])
    assert_bits_equal(sim.u, refjson.array_from_json(snap["u"]), "u")


# ---- the constant-divisor division of the strict kernels equals IEEE `/` -----------------
def test_fastdiv_equals_ieee_division(tmp_path):
    """oracle/fastdiv_check.c: Markstein correction steps vs `/` over ~6 M quotients
    (the divisors of every BASELINE config + random ones, adversarial dividends)."""
    import subprocess
    from pathlib import Path
    src = Path(__file__).resolve().parent.parent / "oracle" / "fastdiv_check.c"
    exe = tmp_path / "fastdiv_check"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), str(src), "-lm"], check=True)
    r = subprocess.run([str(exe), "4000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert " 0 mismatches" in r.stdout


# ---- N4 colour mapping: no golden data in the reference (parity unpinned); the C
#      restatement is cross-checked against an independent numpy-f32 restatement ----------
def _hsl_numpy(hue):
    f = np.float32
    hue = hue.astype(f)
    with np.errstate(invalid="ignore"):
        x = f(1.0) - np.abs(np.fmod(hue / f(60.0), f(2.0)) - f(1.0))
    z, c = np.zeros_like(hue), np.ones_like(hue)
    conds = [hue < 60, hue < 120, hue < 180, hue < 240, hue < 300]
    r = np.select(conds, [c, x, z, z, x], c)
    g = np.select(conds, [x, c, c, x, z], z)
    b = np.select(conds, [z, z, x, c, c], x)

    def u8(ch):
        v = ch * f(255.0)
        return np.where(np.isnan(v) | (v <= 0), 0, np.minimum(v, 255)).astype(np.uint8)
    return np.stack([u8(r), u8(g), u8(b), np.full(hue.shape, 255, np.uint8)], axis=-1)


@pytest.mark.parametrize("color_type", ["pressure", "speed"])
def test_render_restatements_agree(color_type):
    nx, ny = 60, 33
    rng = np.random.default_rng(9)
    kind, bu, bv = po.preset("obstacle", nx, ny)
    o = po.OracleSim(nx, ny, delx=0.1, dely=0.2, delt=0.005, gamma=0.9, reynolds=100.0,
                     sor_absolute_epsilon=1e-3, max_iterations=20, omega=1.7, kind=kind,
                     bu=bu, bv=bv, p=rng.uniform(-3, 3, (nx, ny)), u=rng.uniform(-1, 1, (nx, ny)),
                     v=rng.uniform(-1, 1, (nx, ny)))
    o.run_simulation_tick()
    st = o.state()
    img = o.render_simulation(color_type)
    if color_type == "pressure":
        q, (lo, hi) = np.array(o.p), st.pressure_range
    else:
        q, (lo, hi) = np.sqrt(np.array(o.u) ** 2 + np.array(o.v) ** 2), st.speed_range
    with np.errstate(invalid="ignore", divide="ignore"):
        hue = 240.0 - (q - lo) * 240.0 / (hi - lo)
    want = _hsl_numpy(hue)
    wall = (127, 0, 0, 255) if color_type == "pressure" else (127, 127, 127, 255)
    want[np.array(kind) != 0] = wall
    assert np.array_equal(img, want.transpose(1, 0, 2))
