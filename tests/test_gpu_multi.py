"""Multi-GPU parity (needs >= 2 GPUs): a run whose chains are sharded over 2 ranks, with either
exchange (fused peer stores / NCCL all-gather), is bit-identical to the single-GPU run."""
import hashlib
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(engine, drv, K, nb):
    for _ in range(nb):
        drv.ns_step(K)
    s = engine.state()
    h = hashlib.sha1(engine.dead().tobytes() + engine.live().tobytes()).hexdigest()
    return dict(logZ=s["logZ"], steps=s["total_steps"], evals=s["total_evals"], nmcmc=s["nmcmc"], it=s["iteration"], hash=h)


def _make(rank, world, lanes=16, mix=None, two_lag=True):
    from cpnest_b200.engine import Engine
    from cpnest_b200.proposal import DefaultProposalCycle
    np.random.seed(5)
    e = Engine("gaussian_iid", [[-10, 10]] * 10, nlive=2000, maxmcmc=400, seed=5, device=rank, lanes=lanes,
               rank=rank, world_size=world, mix=mix, two_lag=two_lag)
    e.set_cycle(DefaultProposalCycle().device_cycle())
    e.init_prior()
    return e


def _worker(rank, world, port, comm, out, mix=None):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from cpnest_b200.sharded import ShardedEngine
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    # (the mid-chain snapshots of the two-lag chain-length control travel with the fused exchange only: a run over the NCCL
    # record has the single-lag control, and is compared with a single-GPU run that has it too)
    e = _make(rank, world, mix=mix, two_lag=(comm == "fused"))
    e.set_stream(stream.cuda_stream)
    drv = ShardedEngine(e, comm=comm)
    out[rank] = _run(e, drv, 256, 12)
    drv.close()
    e.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("comm", ["fused", "nccl"])
def test_sharded_run_equals_single_gpu(comm):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    e = _make(0, 1, two_lag=(comm == "fused"))
    ref = _run(e, e, 256, 12)
    e.close()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, comm, out), nprocs=2, join=True)
    assert out[0] == out[1], (out[0], out[1])
    assert out[0] == ref, (out[0], ref)


def test_sharded_mixed_population_equals_single_gpu():
    """A mixed sampler population (MH : slice : HMC = 2 : 1 : 1) sharded over 2 ranks: rank 0 evolves the MH range, rank 1
    the slice and HMC ranges; the result is the single-GPU run's, bit for bit."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mix = (2, 1, 1)
    e = _make(0, 1, mix=mix)
    ref = _run(e, e, 256, 12)
    assert e.state()["chains_kind"] == [128, 64, 64]
    e.close()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, "fused", out, mix), nprocs=2, join=True)
    assert out[0] == out[1], (out[0], out[1])
    assert out[0] == ref, (out[0], ref)
