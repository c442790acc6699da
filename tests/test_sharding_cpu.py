"""Host-side logic of the multi-GPU path, on CPU with the gloo backend (world_size 2):
the chain slices cover a batch exactly once, and the record layout used for the NCCL all-gather
(csrc/api.cu k_pack_shard / k_unpack_shards: [chains][Dp+5] doubles {x.., logL, logP, (steps,acc),
(start,evals), (ne,nc)}) round-trips through all_gather_into_tensor."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cpnest_b200.sharded import shard_range


def test_shard_ranges_partition_the_batch():
    for world in (1, 2, 3, 4, 8):
        for K in (1, 7, 8, 2500, 2501, 20000):
            edges = [shard_range(K, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == K
            for (a0, a1), (b0, b1) in zip(edges[:-1], edges[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_kind_ranges_split_a_batch_and_compose_with_the_shards():
    """cpn_kind_ranges (host-only ABI function): the chains of a batch by sampler kind - exact partition of [0, K), every kind
    present gets a chain, shares within one chain of the requested proportions - and, composed with the rank slices, every
    chain of the batch is evolved exactly once by exactly one (rank, kind) launch whatever the number of GPUs."""
    from cpnest_b200._lib import CpnError
    from cpnest_b200.engine import Engine
    for mix in ((1, 0, 0), (0, 1, 0), (2, 1, 1), (1, 1, 1), (5, 0, 3), (1000, 1, 1), (1, 7, 0)):
        active = sum(m > 0 for m in mix)
        for K in (active, active + 1, 7, 64, 250, 4704, 37632):
            kb = Engine.kind_ranges(mix, K)
            assert kb[0] == 0 and kb[3] == K and all(a <= b for a, b in zip(kb[:-1], kb[1:]))
            n = np.diff(kb)
            for s in range(3):
                assert (n[s] > 0) == (mix[s] > 0)
                if K >= 20 * active and mix[s] > 0 and min(m for m in mix if m > 0) * K >= sum(mix):
                    assert abs(n[s] - K * mix[s] / sum(mix)) <= 1.0 + 1e-9, (mix, K, n)
            for world in (1, 2, 4, 8):
                owner = np.zeros(K, dtype=int)
                for r in range(world):
                    j0, j1 = shard_range(K, r, world)
                    for s in range(3):
                        a, b = max(j0, kb[s]), min(j1, kb[s + 1])
                        if b > a:
                            owner[a:b] += 1
                assert np.all(owner == 1), (mix, K, world)
    assert Engine.kind_ranges((2, 1, 1), 100) == [0, 50, 75, 100]
    assert Engine.kind_ranges((1, 1, 0), 7) == [0, 4, 7, 7]
    for bad_mix, K in (((0, 0, 0), 10), ((1, -1, 0), 10), ((1, 1, 1), 2)):
        with pytest.raises(CpnError):
            Engine.kind_ranges(bad_mix, K)


S_EXTRA = 5      # cpn_shard_record_doubles() - Dp


def pack(j0, j1, Dp, x, logL, logP, steps, acc, start, evals, ne, nc):
    n = j1 - j0
    rec = np.zeros((n, Dp + S_EXTRA))
    rec[:, :Dp] = x[j0:j1]
    rec[:, Dp] = logL[j0:j1]
    rec[:, Dp + 1] = logP[j0:j1]
    ints = np.zeros((n, 6), dtype=np.int32)
    for k, a in enumerate((steps, acc, start, evals, ne, nc)):
        ints[:, k] = a[j0:j1]
    rec[:, Dp + 2:] = ints.view(np.float64)       # (lo, hi) int pairs share a double slot
    return rec


def unpack(rec, Dp):
    ints = np.ascontiguousarray(rec[:, Dp + 2:]).view(np.int32)
    return (rec[:, :Dp], rec[:, Dp], rec[:, Dp + 1]) + tuple(ints[:, k] for k in range(6))


def _worker(rank, world, port, K, Dp, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rs = np.random.RandomState(0)                   # every rank knows the full truth
    x = rs.normal(size=(K, Dp)); logL = rs.normal(size=K); logP = rs.normal(size=K)
    steps = rs.randint(1, 5000, K).astype(np.int32); acc = rs.randint(0, 2000, K).astype(np.int32)
    start = rs.randint(0, 10 ** 6, K).astype(np.int32); evals = rs.randint(0, 5000, K).astype(np.int32)
    ne = rs.randint(0, 10 ** 5, K).astype(np.int32); nc = rs.randint(0, 10 ** 5, K).astype(np.int32)
    j0, j1 = shard_range(K, rank, world)
    # a rank only has ITS slice; poison the rest to prove the gather fills it
    mine = pack(j0, j1, Dp, x, logL, logP, steps, acc, start, evals, ne, nc)
    recv = torch.full((K * (Dp + S_EXTRA),), float("nan"), dtype=torch.float64)
    send = recv[j0 * (Dp + S_EXTRA):j1 * (Dp + S_EXTRA)]
    send.copy_(torch.from_numpy(mine.ravel()))
    dist.all_gather_into_tensor(recv, send.clone())
    got = unpack(recv.numpy().reshape(K, Dp + S_EXTRA), Dp)
    ok = all(np.array_equal(a, b) for a, b in zip(got, (x, logL, logP, steps, acc, start, evals, ne, nc)))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_record_layout_roundtrips_through_gloo_allgather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, 64, 52, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}
