"""Multi-rank host logic on CPU: two processes (gloo), each runs its shard of the phonons (ids == rank mod world),
the integer tallies are summed with an all-reduce exactly as bench.py does over NCCL, and every rank must end up
with the single-shard result bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    from tests import common as T
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = T.load_model(T.case_model("sides_trans"), num_phonons=6_000)
    model.prepare()
    r = T.emu_run(model, 9, shard=rank, num_shards=world)
    # device layout of the tallies: [recorded step][sensor], so a group of steps is one contiguous slice
    energy = torch.from_numpy(np.ascontiguousarray(r["energy"].astype(np.int64).T))
    fixed = torch.from_numpy(np.ascontiguousarray(r["fixed"].transpose(1, 0, 2)))
    steps = torch.tensor([r["drift_steps"]], dtype=torch.int64)
    R = energy.shape[0]
    for lo in range(0, R, 250):  # per group of measurement steps, like the NCCL path
        dist.all_reduce(energy[lo:lo + 250])
        dist.all_reduce(fixed[lo:lo + 250])
    dist.all_reduce(steps)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), energy=energy.numpy().T, fixed=fixed.numpy().transpose(1, 0, 2),
             steps=steps.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_single_shard(tmp_path):
    sys.path.insert(0, ROOT)
    from tests import common as T
    T.emu_lib()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    model = T.load_model(T.case_model("sides_trans"), num_phonons=6_000)
    model.prepare()
    ref = T.emu_run(model, 9)
    for rank in range(2):
        got = np.load(tmp_path / f"rank{rank}.npz")
        assert np.array_equal(got["energy"], ref["energy"])
        assert np.array_equal(got["fixed"], ref["fixed"])
        assert int(got["steps"][0]) == ref["drift_steps"]


def _cuts_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = [0, 899, 935, 971, 999] if rank == 0 else [0, 899, 925, 951, 977, 999]  # e.g. another staging form on this rank
    cuts = bench.agree_on_cuts(mine, 1000, world, rank, "cpu")
    with open(os.path.join(out_dir, f"cuts{rank}.txt"), "w") as f:
        f.write(" ".join(map(str, cuts)))
    dist.barrier()
    dist.destroy_process_group()


def test_ranks_agree_on_the_groups_of_steps(tmp_path):
    """bench.py: every group of measurement steps ends with a collective, so all ranks must use the same cuts - rank 0's -
    even where their own launch windows would differ."""
    import bench
    assert bench.agree_on_cuts([0, 5, 9], 10, 1, 0, "cpu") == [0, 5, 9]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_cuts_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        assert (tmp_path / f"cuts{rank}.txt").read_text() == "0 899 935 971 999"
