"""Ticks on HOST buffers (sb_tick_host, stroemung_b200/pipeline.py) against the oracle.

The reference ticks host arrays in place (src/simulation.rs:324-333); sb_tick_host is that
call for a host-side owner of the fields, and HostPipeline keeps several of them in flight.
Both must give exactly the oracle's tick, whatever the interleaving of the handles."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as po
from stroemung_b200.pipeline import HostPipeline
from stroemung_b200.simulation import SOR_RED_BLACK, SOR_REFERENCE_ORDER, Simulation
from tests.util import assert_bits_equal, oracle_from, random_fields, random_mask, unfinalized

pytestmark = pytest.mark.gpu


def case(nx, ny, seed):
    kind, bu, bv = random_mask(nx, ny, seed, n_blocks=3)
    p, u, v = random_fields(nx, ny, seed)
    return unfinalized(nx, ny, kind, bu, bv, p=p, u=u, v=v)


@pytest.mark.parametrize("mode,omode", [(SOR_REFERENCE_ORDER, po.SOR_REFERENCE_ORDER),
                                        (SOR_RED_BLACK, po.SOR_RED_BLACK)])
def test_tick_host_matches_oracle(mode, omode):
    """three in-place host ticks == three oracle ticks, bit for bit (numpy buffers)"""
    unf = case(96, 150, 5)
    sim = Simulation.try_from(unf, sor_mode=mode)
    o = oracle_from(unf, sor_mode=omode)
    p, u, v = (np.ascontiguousarray(unf["grid"][k], dtype=np.float64).copy() for k in "puv")
    for _ in range(3):
        it, nrm = sim.tick_host(p, u, v)
        oit, onrm = o.run_simulation_tick()
        assert it == oit and abs(nrm - onrm) <= 1e-12 * abs(onrm)
        assert_bits_equal(p, o.p, "p")
        assert_bits_equal(u, o.u, "u")
        assert_bits_equal(v, o.v, "v")
    # out of place: inputs untouched
    p0, u0, v0 = p.copy(), u.copy(), v.copy()
    po_, uo_, vo_ = np.empty_like(p), np.empty_like(p), np.empty_like(p)
    sim.tick_host(p, u, v, po_, uo_, vo_)
    o.run_simulation_tick()
    assert_bits_equal(p, p0, "p input")
    assert_bits_equal(po_, o.p, "p out")
    assert_bits_equal(uo_, o.u, "u out")
    assert_bits_equal(vo_, o.v, "v out")
    sim.close()


def test_tick_host_rejects_null_buffers():
    unf = case(40, 40, 2)
    sim = Simulation.try_from(unf, sor_mode=SOR_RED_BLACK)
    with pytest.raises(Exception):
        sim.tick_host(C.c_void_p(0), C.c_void_p(0), C.c_void_p(0))
    sim.close()


def test_pipeline_three_requests_in_flight():
    """Five independent states through a pipeline of three handles, four ticks each, on pinned
    buffers: every request ends on the oracle's bits whatever handle served which step."""
    nx, ny = 330, 420        # large enough for the pass kernels (streaming + tile kernel)
    kind, bu, bv = random_mask(nx, ny, 11, n_blocks=3)
    # the residual norm a solve has to beat is latched state of a HANDLE
    # (src/simulation.rs:229-237): every handle and every oracle gets the same one
    base = unfinalized(nx, ny, kind, bu, bv, initial_norm_squared=2.5e-3)
    pipe = HostPipeline(lambda: Simulation.try_from(base, sor_mode=SOR_RED_BLACK), depth=3)
    n = nx * ny
    reqs, oracles = [], []
    for r in range(5):
        p, u, v = random_fields(nx, ny, 100 + r)
        bufs = [pipe.alloc() for _ in range(3)]
        views = [np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_double)), shape=(n,)).reshape(nx, ny)
                 for b in bufs]
        for dst, src in zip(views, (p, u, v)):
            dst[...] = src
        reqs.append((bufs, views))
        oracles.append(oracle_from(unfinalized(nx, ny, kind, bu, bv, p=p, u=u, v=v,
                                               initial_norm_squared=2.5e-3),
                                   sor_mode=po.SOR_RED_BLACK))
    results = [[] for _ in reqs]
    for step in range(4):
        futs = [pipe.submit(*bufs) for bufs, _ in reqs]   # 5 requests queued on 3 handles
        for r, f in enumerate(futs):
            results[r].append(f.result())
    for r, o in enumerate(oracles):
        for step in range(4):
            oit, onrm = o.run_simulation_tick()
            it, nrm = results[r][step]
            assert it == oit and abs(nrm - onrm) <= 1e-12 * abs(onrm), (r, step)
        views = reqs[r][1]
        assert_bits_equal(views[0], o.p, f"p of request {r}")
        assert_bits_equal(views[1], o.u, f"u of request {r}")
        assert_bits_equal(views[2], o.v, f"v of request {r}")
    pipe.close()


# ---- `#[derive(Serialize)] Simulation` on the GPU class --------------------------------------

def test_to_json_matches_reference_serialize_snapshot(snapshots):
    """src/simulation.rs:422-447: try_from(presets::empty([5, 7])) serialised ==
    stroemung__simulation__tests__serialize.snap, from the DEVICE state (fields downloaded,
    initial norm latched by sb_create)."""
    g = presets_empty((5, 7))
    unf = {"size": (5, 7), "cell_size": (1.0, 2.0), "delt": 1.4, "gamma": 1.7, "reynolds": 100.0,
           "initial_norm_squared": None, "sor_absolute_epsilon": 0.001, "max_iterations": 100,
           "iterations": 0, "time": 0.0, "omega": 1.7, "grid": g}
    sim = Simulation.try_from(unf, sor_mode=SOR_REFERENCE_ORDER)
    assert sim.to_json() == snapshots["stroemung__simulation__tests__serialize"]["json"]
    sim.close()


def presets_empty(size):
    from stroemung_b200 import presets
    return presets.empty(size)


@pytest.mark.parametrize("sidecar", [False, True])
def test_save_load_round_trip_and_continue(tmp_path, sidecar):
    """tick, save (inline JSON / binary sidecar), load into a new handle, tick both: the same
    bits, and the same as the oracle that never left memory"""
    unf = case(96, 150, 8)
    sim = Simulation.try_from(unf, sor_mode=SOR_RED_BLACK)
    o = oracle_from(unf, sor_mode=po.SOR_RED_BLACK)
    for _ in range(2):
        sim.run_simulation_tick()
        o.run_simulation_tick()
    path = str(tmp_path / "state.json")
    doc = sim.save(path, sidecar=sidecar)
    assert (tmp_path / "state.json.bin").exists() == sidecar
    assert doc["iterations"] == 2 and doc["time"] == o.state().time
    assert doc["initial_norm_squared"] == sim.initial_norm_squared
    sim2 = Simulation.load(path, sor_mode=SOR_RED_BLACK)
    assert np.array_equal(sim2.grid.cell_type, sim.grid.cell_type)
    for s_ in (sim, sim2):
        it, nrm = s_.run_simulation_tick()
        if s_ is sim:
            oit, onrm = o.run_simulation_tick()
        assert it == oit and abs(nrm - onrm) <= 1e-12 * abs(onrm)
        assert_bits_equal(s_.grid.pressure, o.p, "p")
        assert_bits_equal(s_.grid.u, o.u, "u")
        assert_bits_equal(s_.grid.v, o.v, "v")
    assert sim2.to_json() == sim.to_json()
    sim.close()
    sim2.close()


def test_to_json_of_a_device_side_preset():
    """sb_create_preset builds the mask on the device: the velocities of its Inflow cells come
    back through sb_get_boundary_velocities and the document equals the host preset's"""
    from stroemung_b200 import presets, refjson
    size = (100, 20)
    sim = Simulation.from_preset("obstacle", size, (0.1, 0.2), 0.005, 0.9, 100.0, 1e-3, 100, 1.7,
                                 sor_mode=SOR_RED_BLACK)
    g = presets.obstacle(size)
    doc = sim.to_json()
    assert doc["grid"]["cell_type"] == refjson.cells_to_json(g["kind"], g["bu"], g["bv"])
    sim.close()
