"""The sharded path on real GPUs, one process per GPU (skipped on boxes with fewer than two devices):
(a) the product path -- peer-memory communicator, window handles exchanged through CUDA IPC, combine
step fused into the primitives' kernels -- and (b) the same host logic over NCCL collectives (the
baseline bench.py reports next to it). Same checks as test_dist_cpu.py / test_comm_gpu.py, the CUDA
library as the shard-local implementation and the C oracle as the checker."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, fused):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    comm = None
    try:
        from drjit_b200.dist import PeerComm, Sharded
        from tests.dist_body import check_rank
        if fused:
            comm = PeerComm.from_process_group(dist.group.WORLD, dev)
        sh = Sharded(rank=rank, world=world, group=dist.group.WORLD, comm=comm)
        bins = check_rank(sh, rank, world, dev, n)
        # bins must be bit-identical on all ranks (fixed fold order / NCCL guarantee)
        mine = torch.from_numpy(bins).to(dev)
        ref = mine.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(mine, ref)
        torch.cuda.synchronize()
        dist.barrier()
    finally:
        if comm is not None:
            comm.destroy()
        dist.destroy_process_group()


@pytest.mark.parametrize("fused", [True, False], ids=["peer-memory", "nccl"])
@pytest.mark.parametrize("n", [(1 << 22) + 4096])
def test_sharded_primitives_world2(n, fused):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mp.spawn(_worker, args=(2, _free_port(), n, fused), nprocs=2, join=True)
