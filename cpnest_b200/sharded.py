"""Multi-GPU nested sampling: one process per GPU, chains of a batch sharded over the ranks.

Every rank holds a replica of the live set (it is small: nlive * (D+2) doubles) and runs the
selection, the evidence integrator and the merge redundantly and deterministically, exactly as
SURVEY.md section 8e lays out; only the K replacements of a batch cross the NVLink fabric:

  comm='fused' : the replacement staging block lives in torch symmetric memory (peer-mapped over
                 NVLink/NVSwitch); the chain kernel stores every finished chain into all ranks'
                 blocks itself (csrc/mh_kernel.cuh, `out_*[peer]`), so the all-gather overlaps the
                 compute chain by chain and the step needs only a barrier.
  comm='nccl'  : pack -> torch.distributed all_gather_into_tensor (NCCL) -> unpack.  The baseline.

The reference's counterpart is one pipe per sampler process (cpnest/cpnest.py:553-584,
cpnest/NestedSampling.py:331,342).
"""
from __future__ import annotations

import ctypes as C

from . import _lib as L


def shard_range(K, rank, world):
    """Chains [j0, j1) of a K-chain batch evolved by `rank` (same rule as cpn_shard_range)."""
    return (K * rank) // world, (K * (rank + 1)) // world


class ShardedEngine(object):
    """Drives an Engine whose context was created with (rank, world_size) inside a process group."""

    def __init__(self, engine, group=None, comm="fused"):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.e = engine
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.comm = comm
        self.dev = torch.device("cuda", torch.cuda.current_device())
        # The exchange (symmetric-memory barrier / NCCL all-gather) is enqueued by torch, the chain kernel and the merge by the
        # engine: both must use ONE stream, or the barrier could pass before the chain kernel has stored the staging block and
        # the merge could read it before the exchange has finished (replicas would silently diverge).  The engine is moved
        # onto torch's current stream; the legacy default stream (handle 0) cannot be handed over, so a run started on it
        # gets a stream of its own that every exchange is issued on.
        self.stream = torch.cuda.current_stream(self.dev)
        if self.stream.cuda_stream == 0:
            self.stream = torch.cuda.Stream(device=self.dev)
        engine.set_stream(self.stream.cuda_stream)
        self.S = int(engine.lib.cpn_shard_record_doubles(engine.h))
        self._send = self._recv = self._symm = self._hdl = None
        if comm == "fused":
            import torch.distributed._symmetric_memory as symm_mem
            nbytes = int(engine.lib.cpn_staging_bytes(engine.h))
            self._symm = symm_mem.empty((nbytes + 7) // 8, dtype=torch.float64, device=self.dev)
            self._symm.zero_()
            torch.cuda.synchronize(self.dev)
            self._hdl = symm_mem.rendezvous(self._symm, self.group)
            ptrs = (C.c_void_p * self.world)(*[C.c_void_p(int(p)) for p in self._hdl.buffer_ptrs])
            L.check(engine.lib.cpn_set_staging(engine.h, ptrs, self.world))
        elif comm != "nccl":
            raise ValueError("comm must be 'fused' or 'nccl'")

    def _buffers(self, K):
        torch = self.torch
        if self._recv is None or self._recv.numel() < K * self.S:
            self._recv = torch.empty(K * self.S, dtype=torch.float64, device=self.dev)
        return self._recv

    def exchange(self, K):
        """The one exchange step of a batch: after it every rank's staging block holds all K replacements."""
        with self.torch.cuda.stream(self.stream):
            self._exchange(K)

    def _exchange(self, K):
        if self.comm == "fused":
            # the chain kernel has already written into every rank's block; wait for all ranks' kernels
            self._hdl.barrier(channel=0)
            return
        j0, j1 = shard_range(K, self.rank, self.world)
        if (K % self.world) != 0:
            raise ValueError("comm='nccl' needs K divisible by the world size (equal all-gather contributions)")
        recv = self._buffers(K)
        n = j1 - j0
        send = recv[j0 * self.S:(j0 + n) * self.S]           # in place: this rank's slot of the gathered buffer
        L.check(self.e.lib.cpn_pack_shard(self.e.h, K, C.c_void_p(send.data_ptr())))
        self.dist.all_gather_into_tensor(recv[:K * self.S], send, group=self.group)
        L.check(self.e.lib.cpn_unpack_shards(self.e.h, K, C.c_void_p(recv.data_ptr())))

    def ns_step(self, K):
        e = self.e
        e.select_and_accumulate(K)       # redundant on every rank, identical results
        e.evolve_batch(K)                # this rank's slice of the chains
        self.exchange(K)
        e.insert_batch(K)                # redundant, identical

    def run(self, K, max_batches=0, tolerance=0.1, check_every=8):
        """`while condition > tolerance: consume_sample()`; every rank sees the same state."""
        import ctypes
        e = self.e
        s = e.state()
        b0 = s["batches"]
        L.check(e.lib.cpn_set_tolerance(e.h, float(tolerance)))
        n = 0
        while max_batches <= 0 or n < max_batches:
            m = check_every if max_batches <= 0 else min(check_every, max_batches - n)
            for _ in range(m):
                self.ns_step(K)
            n += m
            s = e.state()
            if s["done"] or not (s["condition"] > tolerance):
                break
        return s["batches"] - b0

    def close(self):
        if self.comm == "fused" and self.e.h:
            L.check(self.e.lib.cpn_set_staging(self.e.h, None, 0))
        self._symm = self._hdl = None
