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
