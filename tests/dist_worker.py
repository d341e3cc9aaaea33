"""Worker of tests/test_host_logic.py::test_gather_over_gloo_world_size_2 (CPU, gloo)."""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(sys.argv[1]), int(sys.argv[2])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = importlib.import_module("optimal-control-solvers_b200.host.dist")
    total = 1001                                   # ragged on purpose
    b0, b1 = d.shard_range(total, rank, world)
    J_full = torch.arange(total, dtype=torch.float64) * 0.5
    J = J_full[b0:b1].clone()
    got = d.gather_ragged(J, world, rank, total)
    assert torch.equal(got, J_full)
    # equal shards (the bench's weak-scaling layout)
    Bt = 64
    Jl = torch.full((Bt,), float(rank), dtype=torch.float64)
    xT = torch.full((2, Bt), float(rank) + 0.25, dtype=torch.float64)
    J_all, xT_all = d.gather_results(Jl, xT, world)
    assert J_all.shape == (world * Bt,) and xT_all.shape == (world, 2, Bt)
    for r in range(world):
        assert torch.all(J_all[r * Bt:(r + 1) * Bt] == r) and torch.all(xT_all[r] == r + 0.25)
    # packed, double-buffered gather used by bench.py (synchronous on CPU): to rank 0 (default) and to everybody
    # ("peer" needs CUDA IPC: on CPU tensors it falls back to the NCCL/gloo gather to the root on every rank)
    for mode in ("root", "all", "peer"):
        pg = d.PipelinedGather(2, Bt, world, "cpu", rank=rank, mode=mode)
        assert pg.mode == ("root" if mode == "peer" else mode)
        for step in range(3):
            res = pg.next_buffer()
            res[0] = rank + 10 * step
            res[1:] = rank + 0.5
            pg.gather(res)
            out = pg.out[pg.k]
            if mode == "all" or rank == 0:
                for r in range(world):
                    assert torch.all(out[r, 0] == r + 10 * step) and torch.all(out[r, 1:] == r + 0.5)
            else:
                assert out.shape[0] == 1         # non-root ranks hold nothing
        pg.finish()
    dist.barrier()
    dist.destroy_process_group()
    print("gather ok", rank)


if __name__ == "__main__":
    main()
