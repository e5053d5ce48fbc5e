"""torchrun worker of tests/test_gpu_slabs.py::test_torchrun_two_ranks (and a manual check
at any world size): one process per GPU, slabs connected through CUDA IPC, compared on
rank 0 with a single-GPU run of the same configuration.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node=N --master-addr 127.0.0.1 \\
        --master-port 29541 tests/slab_worker.py [NX NY]
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from stroemung_b200 import multi  # noqa: E402
from stroemung_b200.simulation import SOR_RED_BLACK, Simulation  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    group = multi.TorchGroup(dist)
    nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (640, 384)
    cfg = dict(cell_size=(4.0 / ny, 4.0 / ny), delt=2e-4, gamma=0.9, reynolds=400.0,
               eps=1e-3, max_it=40, omega=1.7)
    args = (nx // 2 + 3, ny // 2, ny / 8.0)   # circle across the middle slab edge

    def make(g, **kw):
        return multi.from_preset(g, "channel_circle", (nx, ny), cfg["cell_size"], cfg["delt"],
                                 cfg["gamma"], cfg["reynolds"], cfg["eps"], cfg["max_it"],
                                 cfg["omega"], preset_args=args, temporal_block=3, **kw)

    sim = make(group, device=local)
    res = [sim.run_simulation_tick() for _ in range(5)]
    fields = {k: multi.gather_field(group, getattr(sim.grid, k)) for k in ("pressure", "u", "v")}
    all_res = group.all_gather(res)
    assert all(r == all_res[0] for r in all_res), "ranks disagree on (iterations, norm)"
    ok = True
    if rank == 0:
        one = multi.ThreadGroup(1).view(0)
        ref = make(one, device=local)
        ref_res = [ref.run_simulation_tick() for _ in range(5)]
        for (a, na), (b, nb) in zip(res, ref_res):
            ok &= a == b and abs(na - nb) <= 1e-12 * max(abs(na), abs(nb))
        for k, v in fields.items():
            same = np.array_equal(v.view(np.uint64), getattr(ref.grid, k).view(np.uint64))
            ok &= bool(same)
            print(f"{k}: bit-identical to the single-GPU run: {same}")
        print("sweeps per tick", [r[0] for r in res], "launches", sim.kernel_launches)
        ref.close()
    group.barrier()
    sim.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(flag[0]) != 1:
        sys.exit(1)
    if rank == 0:
        print("slab_worker ok")


if __name__ == "__main__":
    main()
