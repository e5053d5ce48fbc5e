"""Row-slab multi-GPU runs: one `Simulation` handle per GPU, one process per GPU.

The reference is a single-process solver; this is host plumbing for the build's slab
decomposition (include/stroemung_b200.h "multi-GPU").  Nothing here computes:
the hosts only agree on the partition, pass the connection blobs around and merge the
sparse boundary-velocity tables; halos and reductions move GPU-to-GPU inside the library.

Two interchangeable "groups" carry the host-side all-gather:
  * TorchGroup   -- torch.distributed (NCCL on the GPU box, gloo in the CPU tests);
  * ThreadGroup  -- `world` threads of ONE process, each driving its own handle (the
                    slab tests on a single GPU, and hosts that own several GPUs).
"""
import ctypes as C
import threading

import numpy as np

from . import _capi

HALO = _capi.SLAB_HALO


def slab_range(nx, rank, world):
    """Rows [x_begin, x_end) of slab `rank`: nx rows dealt as evenly as possible, the
    first nx % world slabs one row longer.  Every slab needs >= HALO rows."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(nx, world)
    if base < HALO:
        raise ValueError(f"{nx} rows over {world} slabs leaves fewer than {HALO} rows per slab")
    xb = rank * base + min(rank, rem)
    return xb, xb + base + (1 if rank < rem else 0)


class TorchGroup:
    """Host-side collectives over an initialised torch.distributed process group."""

    def __init__(self, dist, group=None):
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def all_gather(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def barrier(self):
        self.dist.barrier(group=self.group)


class ThreadGroup:
    """`world` threads of one process; make one per run and hand `view(rank)` to thread rank."""

    def __init__(self, world):
        self.world = world
        self._slots = [None] * world
        self._bar = threading.Barrier(world)

    def view(self, rank):
        return _ThreadGroupView(self, rank)


class _ThreadGroupView:
    def __init__(self, parent, rank):
        self._p, self.rank, self.world = parent, rank, parent.world

    def all_gather(self, obj):
        self._p._slots[self.rank] = obj
        self._p._bar.wait()
        out = list(self._p._slots)
        self._p._bar.wait()
        return out

    def barrier(self):
        self._p._bar.wait()


def merge_velocity_tables(tables, x_begin, x_end, halo=HALO):
    """Union of the per-rank sparse tables [(x, y, u, v), ...] (global x), cut down to the
    rows this slab sees: its own plus `halo` rows either side."""
    lo, hi = x_begin - halo, x_end + halo
    seen = {}
    for tab in tables:
        for x, y, u, v in tab:
            if lo <= x < hi:
                seen[(int(x), int(y))] = (float(u), float(v))
    return [(x, y, u, v) for (x, y), (u, v) in sorted(seen.items())]


def local_velocity_table(kind, bu, bv, x_begin):
    """Sparse table of the Inflow / MovingWall cells of this slab's own rows."""
    xs, ys = np.nonzero((kind == _capi.KIND_INFLOW) | (kind == _capi.KIND_MOVING_WALL))
    return [(int(x) + x_begin, int(y), float(bu[x, y]), float(bv[x, y])) for x, y in zip(xs, ys)]


def connect(sim, group):
    """Steps 2 and 3 of the slab protocol: export, all-gather the blobs, connect."""
    blobs = group.all_gather(sim.slab_export())
    sim.slab_connect(blobs)
    return sim


def from_preset(group, preset, size, *args, device=None, **kw):
    """Simulation.from_preset for this rank's slab, connected.  Collective."""
    from .simulation import SOR_RED_BLACK, Simulation
    xb, xe = slab_range(size[0], group.rank, group.world)
    kw.setdefault("sor_mode", SOR_RED_BLACK)
    if group.world > 1:
        kw.update(x_begin=xb, x_end=xe, rank=group.rank, world=group.world)
    if device is not None:
        kw["device"] = device
    sim = Simulation.from_preset(preset, size, *args, **kw)
    return connect(sim, group) if group.world > 1 else sim


def try_from(group, unfinalized, device=None, **kw):
    """Simulation.try_from for this rank's slab.  `unfinalized["grid"]` holds this slab's
    OWN rows of kind / bu / bv / p / u / v ([x_end - x_begin, ny] arrays); "size" is the
    global size.  Collective."""
    from .simulation import SOR_RED_BLACK, Simulation
    nx = unfinalized["size"][0]
    xb, xe = slab_range(nx, group.rank, group.world)
    kw.setdefault("sor_mode", SOR_RED_BLACK)
    if device is not None:
        kw["device"] = device
    if group.world == 1:
        return Simulation.try_from(unfinalized, **kw)
    g = unfinalized["grid"]
    kind = np.asarray(g["kind"])
    assert kind.shape[0] == xe - xb, (kind.shape, xb, xe)
    bu = g.get("bu") if g.get("bu") is not None else np.zeros(kind.shape)
    bv = g.get("bv") if g.get("bv") is not None else np.zeros(kind.shape)
    tables = group.all_gather(local_velocity_table(kind, bu, bv, xb))
    table = merge_velocity_tables(tables, xb, xe)
    kw.update(x_begin=xb, x_end=xe, rank=group.rank, world=group.world)
    sim = Simulation.try_from(unfinalized, velocity_table=table, **kw)
    return connect(sim, group)


def gather_field(group, local):
    """All slabs' rows of one field stacked into the global [nx, ny] array (every rank)."""
    return np.concatenate(group.all_gather(np.ascontiguousarray(local)), axis=0)


def run_threads(world, fn):
    """Run fn(group_view) on `world` threads of this process; returns the results in rank
    order, re-raising the first exception."""
    tg = ThreadGroup(world)
    results, errors = [None] * world, [None] * world

    def body(rank):
        try:
            results[rank] = fn(tg.view(rank))
        except BaseException as e:  # noqa: BLE001 - reported to the caller below
            errors[rank] = e
            tg._bar.abort()

    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    real = [e for e in errors if e is not None and not isinstance(e, threading.BrokenBarrierError)]
    if real or any(errors):
        raise (real or [e for e in errors if e is not None])[0]
    return results


def blob_buffer(blobs):
    """The `world` blobs in rank order as one ctypes byte buffer for sb_slab_connect."""
    n = _capi.SLAB_BLOB_BYTES
    assert all(len(b) == n for b in blobs)
    return (C.c_uint8 * (n * len(blobs))).from_buffer_copy(b"".join(blobs))
