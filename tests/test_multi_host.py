"""Host-side logic of the row-slab (N > 1) path on CPU: the partition, the host groups that
carry blobs / tables (torch.distributed with gloo at world_size 2, and the thread group),
the merge of the sparse boundary-velocity tables and the reassembly of global fields.
No CUDA calls: the device side is covered by tests/test_gpu_slabs.py."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from stroemung_b200 import _capi, multi, presets  # noqa: E402


def test_slab_range_partitions_every_row_once():
    for nx in (20, 33, 100, 8192, 32768, 1000003):
        for world in (1, 2, 3, 4, 8):
            if nx // world < multi.HALO:
                with pytest.raises(ValueError):
                    multi.slab_range(nx, 0, world)
                continue
            ranges = [multi.slab_range(nx, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == nx
            for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
                assert a1 == b0
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= multi.HALO
    with pytest.raises(ValueError):
        multi.slab_range(100, 3, 3)


def test_merge_velocity_tables_keeps_own_and_halo_rows():
    tabs = [[(0, 1, 1.0, 0.0), (5, 2, 0.5, 0.25)], [(30, 4, -1.0, 2.0), (45, 7, 3.0, 3.0)],
            [(30, 4, -1.0, 2.0)]]  # duplicate entry from another rank
    got = multi.merge_velocity_tables(tabs, x_begin=20, x_end=40)
    assert got == [(30, 4, -1.0, 2.0), (45, 7, 3.0, 3.0)]   # rows [10, 50)
    assert multi.merge_velocity_tables(tabs, 0, 10) == [(0, 1, 1.0, 0.0), (5, 2, 0.5, 0.25)]


def test_local_velocity_table_uses_global_x():
    g = presets.simple_inflow((40, 12))
    xb, xe = multi.slab_range(40, 0, 2)
    tab = multi.local_velocity_table(g["kind"][xb:xe], g["bu"][xb:xe], g["bv"][xb:xe], xb)
    assert tab == [(0, y, 1.0, 0.0) for y in range(1, 11)]
    xb, xe = multi.slab_range(40, 1, 2)
    assert multi.local_velocity_table(g["kind"][xb:xe], g["bu"][xb:xe], g["bv"][xb:xe], xb) == []


def test_thread_group_all_gather_and_fields():
    nx, ny, world = 50, 7, 3
    full = np.arange(nx * ny, dtype=np.float64).reshape(nx, ny)

    def body(group):
        xb, xe = multi.slab_range(nx, group.rank, group.world)
        blobs = group.all_gather(bytes([group.rank]) * _capi.SLAB_BLOB_BYTES)
        buf = multi.blob_buffer(blobs)
        assert len(buf) == world * _capi.SLAB_BLOB_BYTES
        assert [buf[r * _capi.SLAB_BLOB_BYTES] for r in range(world)] == list(range(world))
        group.barrier()
        return multi.gather_field(group, full[xb:xe])

    for out in multi.run_threads(world, body):
        assert np.array_equal(out, full)


def test_run_threads_reraises_the_first_error():
    def body(group):
        if group.rank == 1:
            raise KeyError("boom")
        group.barrier()
        return group.rank

    with pytest.raises(KeyError):
        multi.run_threads(2, body)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        group = multi.TorchGroup(dist)
        assert (group.rank, group.world) == (rank, world)
        nx, ny = 46, 9
        g = presets.simple_inflow((nx, ny))
        kind = g["kind"].copy()
        kind[30:34, 3:6] = _capi.KIND_INFLOW       # interior inflow block in slab 1
        bu = g["bu"].copy()
        bu[30:34, 3:6] = 0.5
        xb, xe = multi.slab_range(nx, rank, world)
        tables = group.all_gather(multi.local_velocity_table(kind[xb:xe], bu[xb:xe],
                                                             g["bv"][xb:xe], xb))
        merged = multi.merge_velocity_tables(tables, xb, xe)
        blobs = group.all_gather(bytes([65 + rank]) * _capi.SLAB_BLOB_BYTES)
        field = multi.gather_field(group, kind[xb:xe].astype(np.float64))
        group.barrier()
        q.put((rank, (xb, xe), merged, [b[:1] for b in blobs], field.tolist()))
    finally:
        dist.destroy_process_group()


def test_torch_group_over_gloo_world_size_2():
    """The same host plumbing bench.py uses under torchrun, here over gloo on CPU."""
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nx, ny = 46, 9
    g = presets.simple_inflow((nx, ny))
    kind = g["kind"].copy()
    kind[30:34, 3:6] = _capi.KIND_INFLOW
    inflow_col = [(0, y, 1.0, 0.0) for y in range(1, ny - 1)]
    block = [(x, y, 0.5, 0.0) for x in range(30, 34) for y in range(3, 6)]
    (r0, rng0, tab0, blobs0, f0), (r1, rng1, tab1, blobs1, f1) = got
    assert (r0, r1) == (0, 1) and rng0 == (0, 23) and rng1 == (23, 46)
    # slab 0 sees rows [-10, 33): the inflow column and the block rows 30..32
    assert tab0 == sorted(inflow_col + [e for e in block if e[0] < 33])
    # slab 1 sees rows [13, 56): the whole block, not the inflow column
    assert tab1 == sorted(block)
    assert blobs0 == blobs1 == [b"A", b"B"]
    assert np.array_equal(np.array(f0), kind) and np.array_equal(np.array(f1), kind)
