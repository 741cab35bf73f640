"""A thread-based communicator: `world` shards live in ONE process on ONE GPU, each driven by its own thread.

TEST INFRASTRUCTURE ONLY.  It implements the interface of multiprime_b200.comm.TorchComm with `on_gpu = True`, so a
single-GPU box exercises the DEVICE branches of the sharded path — Hist.export_dev -> all-to-all of device tensors ->
mpb_hist_merge_segments from device pointers, and the in-place all-reduce of the walk's device count vector — that
otherwise only run under NCCL with two or more GPUs.  All shards use the legacy default stream, so device work is ordered
by issue order and the host side by barriers.  (The peer-memory all-reduce is NOT exercised here: kernels of several
shards that wait for each other on one GPU can be serialised by the hardware queues and never finish; its kernel is
tested with the two phases of a round played in turn on one stream, and end to end by the two-GPU NCCL tests.)"""
from __future__ import annotations

import threading

import numpy as np


class _Shared:
    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world


class ThreadComm:
    on_gpu = True

    peer_ok = False

    def __init__(self, shared: _Shared, rank: int, device):
        import torch
        self.torch = torch
        self.sh, self.rank, self.world, self.device = shared, rank, shared.world, device

    # -- plumbing -----------------------------------------------------------------------------------------
    def _exchange(self, obj):
        """everybody deposits, everybody reads everybody's"""
        self.sh.slots[self.rank] = obj
        self.sh.barrier.wait()
        got = list(self.sh.slots)
        self.sh.barrier.wait()
        return got

    def barrier(self):
        self.sh.barrier.wait()

    # -- host collectives -----------------------------------------------------------------------------------
    def allreduce_sum(self, arr):
        arr = np.asarray(arr)
        parts = self._exchange(arr.copy())
        return np.sum(np.stack(parts), axis=0).astype(arr.dtype).reshape(arr.shape)

    def allgather_concat(self, arr):
        parts = self._exchange(np.ascontiguousarray(arr).copy())
        return np.concatenate(parts), np.array([len(p) for p in parts], np.int64)

    def allgather_object(self, obj):
        return self._exchange(obj)

    def allgather_fixed(self, arr):
        return np.stack(self._exchange(np.ascontiguousarray(arr).copy()))

    def alltoall(self, arr, send_counts, recv_counts):
        parts = self._exchange((np.ascontiguousarray(arr).copy(), np.asarray(send_counts)))
        out = []
        for a, sc in parts:
            off = np.concatenate([[0], np.cumsum(sc)])
            out.append(a[off[self.rank]:off[self.rank + 1]])
        res = np.concatenate(out)
        assert [len(o) for o in out] == [int(c) for c in recv_counts]
        return res

    # -- device collectives ---------------------------------------------------------------------------------
    def empty_dev(self, n: int, dtype):
        tdt = {np.dtype(np.uint64): self.torch.int64, np.dtype(np.uint32): self.torch.int32,
               np.dtype(np.int64): self.torch.int64}[np.dtype(dtype)]
        return self.torch.empty(max(1, n), dtype=tdt, device=self.device)

    def alltoall_dev(self, t, send_counts, recv_counts):
        parts = self._exchange_keep((t, np.asarray(send_counts)))
        out = []
        for a, sc in parts:
            off = np.concatenate([[0], np.cumsum(sc)])
            out.append(a[int(off[self.rank]):int(off[self.rank + 1])])
        res = self.torch.cat(out) if out else t[:0]
        self.sh.barrier.wait()                    # nobody frees a tensor another shard is still reading
        assert [len(o) for o in out] == [int(c) for c in recv_counts]
        if res.numel() == 0:
            res = self.torch.empty(1, dtype=t.dtype, device=t.device)
        return res

    def _exchange_keep(self, obj):
        self.sh.slots[self.rank] = obj
        self.sh.barrier.wait()
        return list(self.sh.slots)

    def allreduce_dev_inplace(self, t):
        parts = self._exchange_keep(t)
        total = parts[0].clone()
        for p in parts[1:]:
            total += p
        self.sh.barrier.wait()                    # every shard has its sum before anybody overwrites an input
        t.copy_(total)
        self.sh.barrier.wait()

    def wrap_dev(self, ptr: int, n: int):
        class _Raw:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}
        return self.torch.as_tensor(_Raw(), device=self.device)


def run_shards(world: int, fn):
    """run fn(rank, comm) on `world` threads; returns the results in rank order (re-raises the first failure)"""
    import torch
    shared = _Shared(world)
    out, errs = [None] * world, []

    def body(rank):
        try:
            out[rank] = fn(rank, ThreadComm(shared, rank, torch.device("cuda", 0)))
        except BaseException as exc:              # a dead shard would leave the others at a barrier
            errs.append(exc)
            shared.barrier.abort()

    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    real = [e for e in errs if not isinstance(e, threading.BrokenBarrierError)]
    if real or errs:
        raise (real or errs)[0]
    return out
