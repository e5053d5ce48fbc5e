"""Ticks on HOST buffers, several in flight.

A host-side owner of the fields (the reference keeps `pressure`, `u`, `v` in host arrays,
src/grid/mod.rs:112-125, and `run_simulation_tick` updates them in place,
src/simulation.rs:324-333) pays 3 fields up and 3 fields down per tick: at 8192^2 that is
1.6 GB each way, ~29 ms per direction on PCIe 5 x16 against an 11 ms tick.  One handle does
upload -> tick -> download in series (70 ms).  PCIe is full duplex and the copy engines run
beside the SMs, so `HostPipeline` keeps `depth` handles of the same geometry, each on its own
CUDA stream and driven by its own host thread through `sb_tick_host`: the upload of one
request overlaps the kernels of a second and the download of a third, and throughput is
bounded by the slower PCIe direction instead of the sum.

    pipe = HostPipeline(lambda: Simulation.from_preset(...), depth=3)
    fut = pipe.submit(p, u, v)            # pinned buffers (pipe.alloc()) or numpy arrays
    it, norm = fut.result()               # p, u, v now hold the next time level
    pipe.close()

A request carries the fields only: time, iteration count and the latched
`initial_norm_squared` of the exit rule (src/simulation.rs:229-237, 279) are state of the
handles -- build them with an explicit `initial_norm_squared` when the exit rule matters.
Requests are independent simulations of the same geometry and parameters (ensembles, parameter
sweeps over initial states); a single simulation should stay on the device and use
`Simulation.run_ticks`, which moves nothing.
"""
import ctypes as C
import queue
import threading
from concurrent.futures import Future

from . import _capi


class HostPipeline:
    def __init__(self, make_sim, depth=3):
        """`make_sim()` builds one single-GPU Simulation; it is called `depth` times."""
        assert depth >= 1
        self.sims = [make_sim() for _ in range(depth)]
        self._q = queue.Queue()
        self._pinned = []
        self._threads = [threading.Thread(target=self._worker, args=(s,), daemon=True)
                         for s in self.sims]
        for t in self._threads:
            t.start()

    # page-locked staging buffers of one field -------------------------------------------
    def alloc(self):
        rows, ny = self.sims[0]._local_shape
        p = _capi.lib().sb_host_alloc(rows * ny * 8)
        if not p:
            raise MemoryError("sb_host_alloc failed")
        self._pinned.append(p)
        return C.c_void_p(p)

    @property
    def field_bytes(self):
        rows, ny = self.sims[0]._local_shape
        return rows * ny * 8

    def _worker(self, sim):
        while True:
            job = self._q.get()
            if job is None:
                return
            fut, bufs = job
            if not fut.set_running_or_notify_cancel():
                continue
            try:
                fut.set_result(sim.tick_host(*bufs))
            except BaseException as e:  # noqa: BLE001 -- handed to the caller
                fut.set_exception(e)

    def submit(self, p, u, v, p_out=None, u_out=None, v_out=None):
        """One tick of the state (p, u, v); the returned future yields (sor_iterations,
        norm_squared) once the outputs (default: in place) hold the new state."""
        fut = Future()
        self._q.put((fut, (p, u, v, p_out, u_out, v_out)))
        return fut

    def close(self):
        for _ in self._threads:
            self._q.put(None)
        for t in self._threads:
            t.join()
        for s in self.sims:
            s.close()
        for p in self._pinned:
            _capi.lib().sb_host_free(C.c_void_p(p))
        self._pinned = []
        self.sims = []
