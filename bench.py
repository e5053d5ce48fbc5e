#!/usr/bin/env python3
"""bench.py -- the headline benchmark of stroemung_b200 (contract: see README / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c5|c1|c2|c3|c4] [--size NX NY] [--mode rb|lex] [--tblock T]

One "step" is one full simulation tick (`Simulation::run_simulation_tick`,
/root/reference/src/simulation.rs:324-333: velocity BC, F/G, RHS, SOR solve incl. the
residual norm after every sweep, velocity update, ranges) on synthetic input.

Default workload (N = 1): BASELINE.json config 5 -- the SOR-dominated full tick on an
8192 x 8192 f64 channel grid (simple_inflow layout, fields at rest), eps = 0 and
max_iterations = 100 so that every tick runs exactly 100 sweeps; for
N GPUs the grid is (8192 N) x 8192, split into N row slabs (weak scaling).  The working
set (7 f64 arrays + flags = 3.8 GB per GPU) is far larger than the 126 MB L2, so no L2
flush is needed between timed steps.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput; `e2e` is the same
metric through the C ABI with HOST buffers: every step uploads p, u, v from pinned host
memory, ticks, and downloads p, u, v again.

--impl reference times the reference's own CPU algorithm (the C oracle, single-threaded
like the Rust original) on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "Mcell-steps/s (full step incl. SOR)"
UNIT = "Mcell-steps/s"
SOR_BYTES_PER_CELL_SWEEP = 25.0   # read p, rhs, flag; write p
TICK_FIXED_BYTES_PER_CELL = 81.0  # F/G+RHS 40 + velocity update/ranges 41
# dram__bytes_read.sum + dram__bytes_write.sum per SOR pass (the streaming kernel plus the
# tile kernel on what is left), from the committed `ncu --set full` captures: see
# profiles/ncu_traffic.json (key "<mode>-T<T>-<rows>x<ny>"); None until a capture exists


def ncu_traffic(key):
    f = ROOT / "profiles" / "ncu_traffic.json"
    if not f.exists():
        return None
    e = json.loads(f.read_text()).get(key)
    return e["bytes_per_pass"] if e else None



def workload(name, n_gpus, size=None):
    """-> dict(preset, preset_args, size, scalar parameters) of one BASELINE config."""
    wl = _workload(name, n_gpus, size)
    wl["key"] = name
    # only c5 grows with N ((8192 N) x 8192); every other workload is a fixed grid cut into N slabs
    wl["scaling"] = "weak" if (name == "c5" and size is None) else "strong"
    return wl


def _workload(name, n_gpus, size=None):
    if name in ("c5", "c5s"):   # SOR-dominated scaling sweep: channel, 100 fixed sweeps per tick
        # c5: weak -- (8192 N) x 8192 on N GPUs; c5s: strong -- 32768^2 (or --size) whatever N
        nx, ny = size or ((8192 * n_gpus, 8192) if name == "c5" else (32768, 32768))
        return dict(name="c5-sor-dominated-channel", preset="simple_inflow", preset_args=(),
                    size=(nx, ny), cell_size=(10.0 / ny, 10.0 / ny), delt=2e-4, gamma=0.9,
                    reynolds=100.0, eps=0.0, max_iterations=100, omega=1.7)
    if name == "c1":   # default obstacle preset, CLI defaults of the reference (src/args.rs)
        nx, ny = size or (100, 20)
        return dict(name="c1-obstacle-default", preset="obstacle", preset_args=(),
                    size=(nx, ny), cell_size=(0.1, 0.2), delt=0.005, gamma=0.9,
                    reynolds=100.0, eps=1e-3, max_iterations=100, omega=1.7)
    if name == "c2":   # lid-driven cavity (extension kind; not expressible in the reference)
        nx, ny = size or (1024, 1024)
        return dict(name="c2-lid-driven-cavity", preset="cavity", preset_args=(1.0,),
                    size=(nx, ny), cell_size=(1.0 / (nx - 2), 1.0 / (ny - 2)), delt=1e-4,
                    gamma=0.9, reynolds=1000.0, eps=1e-3, max_iterations=1000, omega=1.7)
    if name == "c3":   # Karman vortex street past a circle
        nx, ny = size or (8192, 2048)
        return dict(name="c3-karman-circle", preset="channel_circle",
                    preset_args=(nx // 8, ny // 2, ny / 16.0), size=(nx, ny),
                    cell_size=(4.1 / (ny - 2), 4.1 / (ny - 2)), delt=2e-4, gamma=0.9,
                    reynolds=400.0, eps=1e-3, max_iterations=100, omega=1.7)
    if name == "c4":   # backward-facing step
        nx, ny = size or (16384, 4096)
        return dict(name="c4-backward-step", preset="backward_step",
                    preset_args=(nx // 4, ny // 2), size=(nx, ny),
                    cell_size=(7.5 / (ny - 2), 7.5 / (ny - 2)), delt=1e-4, gamma=0.9,
                    reynolds=200.0, eps=1e-3, max_iterations=100, omega=1.7)
    raise SystemExit(f"unknown workload {name}")


# ---- clocks ---------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region (the quantities of
    the B200_PROFILING.md clocks line), read through NVML from a background thread every
    ~5 ms -- the timed region of a 20-tick run is a fraction of a second, too short for
    `nvidia-smi -lms`.  Only samples taken between begin() and end() are reported."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
               "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.samples = []
        self.active = False
        self.stop_flag = False
        self.thread = None
        self.smax = None

    def _loop(self, nv, h):
        while not self.stop_flag:
            if self.active:
                try:
                    sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                    pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                    try:
                        rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:
                        rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    self.samples.append((sm, pw, rs))
                except Exception:
                    pass
            time.sleep(0.005)

    def start(self):
        try:
            import threading

            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu_index
            if vis:
                idx = int(vis.split(",")[self.gpu_index])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.smax = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._loop, args=(nv, h), daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None

    def begin(self):
        self.active = True

    def end(self):
        self.active = False

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.smax, "reasons": [], "samples": 0}
        if self.thread is None:
            return out
        self.stop_flag = True
        self.thread.join(timeout=2)
        if self.samples:
            reasons = sorted(name for name, bit in self.REASONS.items()
                             if any(r & bit for _, _, r in self.samples))
            out.update(sm_mhz=statistics.median(x[0] for x in self.samples),
                       samples=len(self.samples), reasons=reasons,
                       power_w_max=max(x[1] for x in self.samples))
        return out


def measured_peak_gbs():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---- reference arm / CPU baseline ------------------------------------------------------------
def cpu_sample(wl, ticks, sample=(2048, 2048)):
    """Time `ticks` full ticks of the C oracle (reference-order SOR, 1 thread -- the
    reference is single-threaded safe Rust) on a bounded sample of the workload: the same
    preset, parameters and sweeps per tick on a smaller grid.  Returns Mcell-steps/s."""
    from oracle import pyoracle as po
    from stroemung_b200 import presets
    nx, ny = min(sample[0], wl["size"][0]), min(sample[1], wl["size"][1])
    sub = workload(wl["key"], 1, (nx, ny))  # same shape rules, smaller grid
    a = sub["preset_args"]
    if sub["preset"] == "channel_circle":
        g = presets.channel_circle((nx, ny), int(a[0]), int(a[1]), float(a[2]))
    elif sub["preset"] == "backward_step":
        g = presets.backward_step((nx, ny), int(a[0]), int(a[1]))
    elif sub["preset"] == "cavity":
        g = presets.cavity((nx, ny), float(a[0]))
    else:
        g = getattr(presets, sub["preset"])((nx, ny))
    kind, bu, bv = g["kind"], g["bu"], g["bv"]
    o = po.OracleSim(nx, ny, delx=wl["cell_size"][0], dely=wl["cell_size"][1], delt=wl["delt"],
                     gamma=wl["gamma"], reynolds=wl["reynolds"],
                     sor_absolute_epsilon=wl["eps"], max_iterations=wl["max_iterations"],
                     omega=wl["omega"], kind=kind, bu=bu, bv=bv)
    # one thread pinned to one core (the reference is single-threaded)
    pinned = None
    try:
        old = os.sched_getaffinity(0)
        core = max(old)
        os.sched_setaffinity(0, {core})
        pinned = core
    except (AttributeError, OSError):
        old = None
    try:
        t0 = time.perf_counter()
        sweeps = 0
        for _ in range(ticks):
            it, _ = o.run_simulation_tick()
            sweeps += it
        dt = time.perf_counter() - t0
    finally:
        if old is not None:
            os.sched_setaffinity(0, old)
    return {"value": nx * ny * ticks / dt / 1e6, "seconds": dt, "grid": [nx, ny],
            "ticks": ticks, "sweeps": sweeps, "pinned_core": pinned}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = workload(args.workload, args.gpus, args.size)
    # the real grid whenever it is <= 2048^2 cells, else a 2048^2 sample of the same preset
    sample = (2048, 2048) if wl["size"][0] * wl["size"][1] > 2048 * 2048 else wl["size"]
    cpu_sample(wl, max(args.warmup, 0) and 1, sample=(256, 256))  # warm the code / caches
    # bounded: a 2048^2 tick of 100 sweeps takes ~6 s on one core, 20 steps ~2 minutes
    r = cpu_sample(wl, args.steps, sample)
    sample_txt = (f"{r['grid'][0]}x{r['grid'][1]} grid of the same preset and parameters"
                  f"{' (a sample: per-cell rate)' if list(r['grid']) != list(wl['size']) else ''}, "
                  f"{r['ticks']} ticks, {r['sweeps']} lexicographic SOR sweeps, {r['seconds']:.1f} s, "
                  f"C oracle (port of the Rust reference), 1 thread pinned to core "
                  f"{r['pinned_core']}")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * r["seconds"] / max(r["ticks"], 1), "higher_is_better": True,
        "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "grid": list(wl["size"]), "sample": sample_txt,
                   "sweeps_per_tick": r["sweeps"] / max(r["ticks"], 1)},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": sample_txt},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ---- verification on the bench grid ----------------------------------------------------------
def host_preset(wl):
    from stroemung_b200 import presets
    nx, ny = wl["size"]
    a = wl["preset_args"]
    if wl["preset"] == "channel_circle":
        return presets.channel_circle((nx, ny), int(a[0]), int(a[1]), float(a[2]))
    if wl["preset"] == "backward_step":
        return presets.backward_step((nx, ny), int(a[0]), int(a[1]))
    if wl["preset"] == "cavity":
        return presets.cavity((nx, ny), float(a[0]))
    return getattr(presets, wl["preset"])((nx, ny))


def verify_on_bench_grid(wl, mode, tblock, device, sweeps=8, tick_sweeps=4):
    """The kernels, plan and grid of the timed region against the CPU oracle (the checker,
    never the thing measured): from a seeded random state (p, u, v ~ U(-0.5, 0.5), splitmix-
    seeded numpy generator) `sweeps` SOR sweeps and one full tick capped at `tick_sweeps`
    sweeps, on the bench grid with the bench's mask, parameters and temporal block.
    p after the sweeps and p, u, v after the tick must be bit-identical (red-black: to the
    oracle's red-black restatement; lex: to the reference order), norms within 1e-12."""
    from oracle import pyoracle as po
    from stroemung_b200.simulation import SOR_RED_BLACK, SOR_REFERENCE_ORDER, Simulation
    t0 = time.perf_counter()
    nx, ny = wl["size"]
    rng = np.random.default_rng(0x5EED5EED)
    p, u, v = (rng.uniform(-0.5, 0.5, (nx, ny)) for _ in range(3))
    g = host_preset(wl)
    prm = dict(delx=wl["cell_size"][0], dely=wl["cell_size"][1], delt=wl["delt"],
               gamma=wl["gamma"], reynolds=wl["reynolds"], sor_absolute_epsilon=wl["eps"],
               max_iterations=tick_sweeps, omega=wl["omega"])
    rb = mode == "rb"
    # initial_norm_squared = 0: the tick runs all `tick_sweeps` sweeps (with the norm of the
    # random state latched, the exit rule of src/simulation.rs:279 would fire after one)
    o = po.OracleSim(nx, ny, kind=g["kind"], bu=g["bu"], bv=g["bv"], p=p, u=u, v=v,
                     initial_norm_squared=0.0,
                     sor_mode=po.SOR_RED_BLACK if rb else po.SOR_REFERENCE_ORDER, **prm)
    unf = {"size": (nx, ny), "cell_size": wl["cell_size"], "delt": wl["delt"],
           "gamma": wl["gamma"], "reynolds": wl["reynolds"], "sor_absolute_epsilon": wl["eps"],
           "max_iterations": tick_sweeps, "omega": wl["omega"], "initial_norm_squared": 0.0,
           "grid": {"p": p, "u": u, "v": v, "kind": g["kind"], "bu": g["bu"], "bv": g["bv"]}}
    sim = Simulation.try_from(unf, sor_mode=SOR_RED_BLACK if rb else SOR_REFERENCE_ORDER,
                              temporal_block=tblock, device=device)
    del p, u, v

    def same(a, b):
        return bool(np.array_equal(np.ascontiguousarray(a).view(np.uint64),
                                   np.ascontiguousarray(b).view(np.uint64)))

    # norms: the oracle folds the squared residuals of all cells sequentially like the
    # reference (src/simulation.rs:216-227), the GPU sums a tree; the sequential sum of N
    # positive terms carries a rounding error of ~sqrt(N) ulp (9e-13 at 8192^2)
    rtol = max(1e-12, 8.0 * np.sqrt(float(nx) * ny) * 2.0 ** -53)
    worst = [0.0]

    def close(a, b):
        if a != b:
            worst[0] = max(worst[0], abs(a - b) / max(abs(a), abs(b)))
        return a == b or abs(a - b) <= rtol * max(abs(a), abs(b))

    bad = []
    norms = sim.sor_sweeps(sweeps)
    for k in range(sweeps):
        o.sor_sweep()
        if not close(norms[k], o.calculate_norm_squared()):
            bad.append(f"norm of sweep {k}")
    if not same(sim.grid.pressure, o.p):
        bad.append("p after sweeps")
    it, nrm = sim.run_simulation_tick()
    oit, onrm = o.run_simulation_tick()
    if it != oit:
        bad.append(f"tick iterations {it} vs {oit}")
    if not close(nrm, onrm):
        bad.append(f"tick norm {nrm!r} vs {onrm!r}")
    for name, a, b in (("p", sim.grid.pressure, o.p), ("u", sim.grid.u, o.u),
                       ("v", sim.grid.v, o.v)):
        if not same(a, b):
            bad.append(f"{name} after tick")
    plan = list(sim.rb_plan) if rb else None
    path = sim.sor_path[0] if rb else None
    sim.close()
    return {"result": "bit-exact" if not bad else "MISMATCH: " + ", ".join(bad),
            "grid": [nx, ny], "sweeps": sweeps, "tick_sweeps": it,
            "fields": "p after sweeps; p, u, v after one tick (random initial p, u, v)",
            "norm_rtol": rtol, "norm_rel_diff_max": worst[0], "against": "oracle/stroemung_oracle.c "
            + ("red-black restatement" if rb else "reference order"),
            "rb_plan": plan, "sor_path": path, "seconds": time.perf_counter() - t0}


def _e2e_pipelined(sim, depth, args, fields, L, multi, group, wl, ext):
    """N = 1 end-to-end leg: `depth` requests in flight through HostPipeline (sb_tick_host).
    -> (seconds, steps in them, steps per request, description)"""
    import threading
    import time

    from stroemung_b200.pipeline import HostPipeline
    first = [sim]

    def make_sim():
        if first:
            return first.pop()
        return multi.from_preset(group, wl["preset"], wl["size"], wl["cell_size"],
                                 wl["delt"], wl["gamma"], wl["reynolds"], wl["eps"],
                                 wl["max_iterations"], wl["omega"],
                                 preset_args=wl["preset_args"], **ext)
    pipe = HostPipeline(make_sim, depth=depth)
    req = [[pipe.alloc() for _ in range(3)] for _ in range(depth)]
    for bufs in req:
        for hbuf, fld in zip(bufs, fields):
            sim._check(L.sb_download(sim._h, fld, hbuf))
    # enough steps per request for the requests to fall out of lockstep and for the fill (first
    # uploads with nothing to overlap) and drain (last downloads) to amortise
    e2e_steps = max(2, min(args.steps, 10))
    for bufs in req:                      # one untimed step per request (first-touch)
        pipe.submit(*bufs).result()

    def drive(bufs):
        for _ in range(e2e_steps):
            pipe.submit(*bufs).result()   # H2D x3, tick, D2H x3 inside sb_tick_host
    drivers = [threading.Thread(target=drive, args=(bufs,)) for bufs in req]
    t0 = time.perf_counter()
    for t in drivers:
        t.start()
    for t in drivers:
        t.join()                          # every result() has synchronised its stream
    dt_e2e = time.perf_counter() - t0
    e2e_total = depth * e2e_steps
    extra_sims = [s_ for s_ in pipe.sims if s_ is not sim]
    pipe.sims = []                        # `sim` is closed below, the others here
    pipe.close()
    for s_ in extra_sims:
        s_.close()
    e2e_how = (f"{depth} requests in flight (HostPipeline: {depth} handles, streams and host "
               f"threads, sb_tick_host), {e2e_steps} steps each; host clock around all of them")
    return dt_e2e, e2e_total, e2e_steps, e2e_how


# ---- our arm -----------------------------------------------------------------------------
def run_ours(args):
    from stroemung_b200 import _capi
    from stroemung_b200.simulation import SOR_RED_BLACK, SOR_REFERENCE_ORDER, Simulation

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        # NCCL prints its version banner on stdout at the first communicator: keep stdout
        # clean for the ONE JSON line by pointing fd 1 at stderr until NCCL is up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    n_gpus = world
    wl = workload(args.workload, n_gpus, args.size)
    nx, ny = wl["size"]
    mode = SOR_RED_BLACK if args.mode == "rb" else SOR_REFERENCE_ORDER
    ext = dict(sor_mode=mode, temporal_block=args.tblock, device=local_rank)
    if args.tau > 0:
        ext["tau"] = args.tau     # extension A9: adaptive time step (max |u|, |v| reductions)
    from stroemung_b200 import multi
    # N > 1: one row slab per rank, connected GPU-to-GPU (CUDA IPC over NVLink); the host
    # group only carries the connection blobs (stroemung_b200/multi.py)
    group = multi.TorchGroup(dist) if world > 1 else multi.ThreadGroup(1).view(0)
    sim = multi.from_preset(group, wl["preset"], wl["size"], wl["cell_size"], wl["delt"],
                            wl["gamma"], wl["reynolds"], wl["eps"], wl["max_iterations"],
                            wl["omega"], preset_args=wl["preset_args"], **ext)
    L = _capi.lib()

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # -- warm-up --------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    sweeps = []
    for _ in range(args.warmup):
        it, _ = sim.run_simulation_tick()
    # -- timed region: device-resident ticks -----------------------------------------------
    launches0 = sim.kernel_launches
    sim.profile_enable(True)
    barrier()
    sampler.begin()
    sim.timer_begin()                         # CUDA event on the stream the kernels run on
    sor_ms = 0.0
    stage_ms = np.zeros(4)
    for _ in range(args.steps):
        it, nrm = sim.run_simulation_tick()   # synchronises the handle's stream at its end
        sweeps.append(it)
        sor_ms += sim.last_sor_ms
        stage_ms += sim.last_stage_ms
    dt = sim.timer_end() * 1e-3               # event + synchronize
    sampler.end()
    barrier()
    clocks = sampler.stop()
    pass_ms = sim.profile_read()
    rb_plan = list(sim.rb_plan) if args.mode == "rb" else None
    sor_path, sor_ctas = sim.sor_path if args.mode == "rb" else (0, 0)
    sim.profile_enable(False)
    launches = sim.kernel_launches - launches0
    dt, sor_ms = max_over_ranks(dt), max_over_ranks(sor_ms)
    if dist is not None:
        import torch
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt[0])
    cells = nx * ny
    value = cells * args.steps / dt / 1e6
    total_sweeps = sum(sweeps)

    # -- end-to-end: host buffers in, host buffers out, every step --------------------------
    # Every step uploads p, u, v from pinned host memory, ticks, downloads p, u, v.  N = 1:
    # through stroemung_b200.pipeline.HostPipeline -- `depth` handles of this geometry, each
    # with its own stream and host thread calling sb_tick_host, so the upload of one request
    # overlaps the kernels of another and the download of a third (PCIe is full duplex); the
    # requests are `depth` independent simulations, each stepped `e2e_steps` times in series
    # (a step's input is the previous output of the same request).  N > 1: one slab handle per
    # rank, upload -> halo sync -> tick -> download in series.
    rows = sim._local_shape[0]
    nbytes = rows * ny * 8
    fields = (_capi.FIELD_P, _capi.FIELD_U, _capi.FIELD_V)
    sim_bytes = rows * ny * 57
    depth = 1 if world > 1 or args.e2e_depth == 1 or 3 * sim_bytes > 60e9 else args.e2e_depth
    e2e_done = False
    if depth > 1:
        try:
            dt_e2e, e2e_total, e2e_steps, e2e_how = _e2e_pipelined(
                sim, depth, args, fields, L, multi, group, wl, ext)
            e2e_done = True
        except (MemoryError, _capi.SbError) as e:   # e.g. not enough page-locked memory
            print(f"[bench] pipelined end-to-end leg unavailable ({e}); serial", file=sys.stderr)
            depth = 1
    if not e2e_done:
        host = [C.c_void_p(L.sb_host_alloc(nbytes)) for _ in range(3)]
        for hbuf, fld in zip(host, fields):
            sim._check(L.sb_download(sim._h, fld, hbuf))
        e2e_steps = max(1, min(args.steps, 3))
        barrier()
        sim.timer_begin()
        for _ in range(e2e_steps):
            for hbuf, fld in zip(host, fields):
                sim._check(L.sb_upload(sim._h, fld, hbuf))      # H2D from pinned host memory
            if world > 1:
                sim.slab_sync_halos()                           # uploaded rows -> neighbours' halos
            sim.run_simulation_tick()
            for hbuf, fld in zip(host, fields):
                sim._check(L.sb_download(sim._h, fld, hbuf))    # D2H of the step's result
        dt_e2e = max_over_ranks(sim.timer_end() * 1e-3)
        barrier()
        for hbuf in host:
            L.sb_host_free(hbuf)
        e2e_total = e2e_steps
        e2e_how = "one handle per rank: upload, tick, download in series; CUDA events, max over ranks"
    e2e_value = cells * e2e_total / dt_e2e / 1e6

    # -- roofline of the dominant kernel (the SOR pass) -------------------------------------
    peak, peak_src = measured_peak_gbs()
    T = (sim.temporal_block or 4) if args.mode == "rb" else 1  # 0 = the library default (4)
    work = [m for m in pass_ms if m > 0.2 * (max(pass_ms) if len(pass_ms) else 1.0)]
    local_cells = rows * ny
    roof = None
    if work:
        avg_ms = sum(work) / len(work)
        sweeps_per_launch = total_sweeps / len(work)
        alg_bytes = SOR_BYTES_PER_CELL_SWEEP * local_cells * sweeps_per_launch
        achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
        key = f"{args.mode}-T{T}-{rows}x{ny}"
        traffic = ncu_traffic(key)
        roof = {"bound": "hbm",
                "kernel": "sor_lex_kernel" if args.mode != "rb" else
                (f"SOR pass = sor_rb_stream_kernel<{T}> on {rb_plan[1]} work items + "
                 f"sor_rb_kernel on {rb_plan[0]} tiles") if sor_path == 0 else
                "whole solve = sor_small_kernel (one SM)" if sor_path == 1 else
                f"whole solve = {'sor_mid_reg_kernel' if sor_path == 3 else 'sor_mid_kernel'}, "
                f"grid resident in the shared memory of {sor_ctas} SMs (cooperative launch)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic,
                "dram_gbs": traffic / (avg_ms * 1e-3) / 1e9 if traffic else None,
                "dram_frac": traffic / (avg_ms * 1e-3) / 1e9 / peak if traffic else None,
                "peak_source": peak_src, "launches_timed": len(work),
                "avg_launch_ms": avg_ms, "sweeps_per_launch": sweeps_per_launch,
                "launch_ms_min_median_max": [min(work), statistics.median(work), max(work)],
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "achieved = 25 B per cell-sweep x cells x sweeps per pass / pass time "
                        "(CUDA events around every pass of the timed region); a pass fuses T "
                        "sweeps, so it moves ~25 B per cell once (traffic, dram_gbs) and the "
                        "algorithmic figure may exceed the HBM peak"}
    barrier()
    sim.close()
    if rank != 0:
        dist.destroy_process_group()
        return 0

    # -- CPU baseline beside it (N = 1 only): bounded sample of the same workload -----------
    cpu = None
    if n_gpus == 1 and not args.no_cpu:
        sample = (2048, 2048) if cells > 2048 * 2048 else wl["size"]
        cells_s = min(sample[0], nx) * min(sample[1], ny)
        r = cpu_sample(wl, 5 if cells_s * wl["max_iterations"] >= 2e7 else min(args.steps, 20),
                       sample)
        cpu = {"value": r["value"], "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{r['grid'][0]}x{r['grid'][1]} grid of the same preset/parameters, "
                         f"{r['ticks']} ticks, {r['sweeps']} lexicographic sweeps, "
                         f"{r['seconds']:.1f} s, C oracle (port of the Rust reference), 1 thread "
                         f"pinned to core {r['pinned_core']}"}
    # -- verification of the timed kernels on the bench grid (N = 1; slabs: tests/) ----------
    verify = None
    if n_gpus == 1 and not args.no_verify:
        if cells <= 8192 * 8192:
            verify = verify_on_bench_grid(wl, args.mode, args.tblock, local_rank)
        else:
            verify = {"result": "skipped", "why": "grid larger than 8192^2: the CPU oracle would "
                      "need minutes; the same plan shapes are verified in tests/test_gpu_scale.py"}

    ws_mb = local_cells * 57 / 1e6   # 7 f64 arrays + flags
    kfix = TICK_FIXED_BYTES_PER_CELL
    k_avg = total_sweeps / max(args.steps, 1)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "grid": [nx, ny], "sor_mode": args.mode,
                   "initial_state": "fields at rest (p = u = v = 0; the inflow starts the flow); "
                                    "f64 arithmetic time does not depend on the values, and "
                                    "`verify` below runs the same kernels from a random state",
                   "temporal_block": T, "sweeps_per_tick": k_avg, "tau": args.tau,
                   "slabs": f"{n_gpus} row slab(s) along x",
                   "rb_plan": {"tile_kernel_tiles": rb_plan[0], "stream_items": rb_plan[1]}
                   if rb_plan and sor_path == 0 else None,
                   "sor_path": ["pass kernels", "sor_small (one launch, one SM)",
                                f"sor_mid (one launch, {sor_ctas} SMs)",
                                f"sor_mid_reg (one launch, {sor_ctas} SMs)"][sor_path],
                   "l2": (f"working set {ws_mb / 1e3:.1f} GB per GPU >> 126 MB L2, no flush needed"
                          if ws_mb > 1000 else
                          f"working set {ws_mb:.0f} MB per GPU: not flushed between ticks (a "
                          f"tick rewrites every array; secondary workload, not the headline)")},
        "sor": {"gcell_sweeps_per_s": cells * total_sweeps / (sor_ms * 1e-3) / 1e9
                if sor_ms else None,
                "algorithmic_gbs": SOR_BYTES_PER_CELL_SWEEP * cells * total_sweeps /
                (sor_ms * 1e-3) / 1e9 if sor_ms else None,
                "ms_per_tick": sor_ms / max(args.steps, 1)},
        "stage_ms_per_tick": dict(zip(("velocity_bc", "fg_rhs", "sor", "set_u_and_v"),
                                      (float(x) / max(args.steps, 1) for x in stage_ms))),
        "tick_roofline": {"bytes_per_cell": kfix + SOR_BYTES_PER_CELL_SWEEP * k_avg,
                          "roofline_mcell_steps_per_s": n_gpus * peak * 1e9 /
                          (kfix + SOR_BYTES_PER_CELL_SWEEP * k_avg) / 1e6,
                          "frac": value / (n_gpus * peak * 1e9 /
                                           (kfix + SOR_BYTES_PER_CELL_SWEEP * k_avg) / 1e6)},
        "roofline": roof, "cpu_baseline": cpu, "verify": verify, "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 3 * nbytes * n_gpus,
                "d2h_bytes_per_step": 3 * nbytes * n_gpus, "steps": e2e_total,
                "in_flight": depth, "how": e2e_how},
        "gpu_launches": launches,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5")
    ap.add_argument("--size", type=int, nargs=2, default=None)
    ap.add_argument("--mode", default="rb", choices=["rb", "lex"])
    ap.add_argument("--tblock", type=int, default=0)
    ap.add_argument("--tau", type=float, default=0.0,
                    help="> 0: adaptive time step (extension A9; the reference has none)")
    ap.add_argument("--e2e-depth", type=int, default=3,
                    help="requests in flight in the end-to-end leg at N = 1 (1 = serial)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    args = ap.parse_args()
    if args.size is not None:
        args.size = tuple(args.size)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
