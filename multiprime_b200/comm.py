"""Sequence-shard communicator: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo in the
CPU tests).  All collectives of the hot path are here: an integer all-reduce of coverage counts per scan round, an
all-reduce of gap counts / entropy bounds and an all-gather of compact haplotype entries per window batch."""
from __future__ import annotations

import numpy as np


class NoComm:
    world = 1
    rank = 0

    def allreduce_sum(self, arr):
        return arr

    def allgather_concat(self, arr):
        return arr, np.array([len(arr)], np.int64)

    def allgather_object(self, obj):
        return [obj]

    def allgather_fixed(self, arr):
        return np.asarray(arr)[None]

    def alltoall(self, arr, send_counts, recv_counts):
        return arr

    def barrier(self):
        pass


class TorchComm:
    def __init__(self, device=None, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" \
                else torch.device("cpu")
        self.device = device
        self.peer_key = ("torch", id(group) if group is not None else 0)

    def _to(self, arr):
        t = self.torch.from_numpy(np.ascontiguousarray(arr))
        return t.to(self.device) if self.device.type != "cpu" else t.clone()

    def allreduce_sum(self, arr):
        """element-wise sum over ranks of an int64 / float64 array"""
        arr = np.asarray(arr)
        kind = np.float64 if arr.dtype.kind == "f" else np.int64
        t = self._to(arr.astype(kind))
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy().astype(arr.dtype).reshape(arr.shape)

    def allgather_concat(self, arr):
        """concatenation over ranks of 1-d arrays of different lengths -> (all, lengths per rank)"""
        arr = np.ascontiguousarray(arr)
        n = self.torch.tensor([arr.shape[0]], dtype=self.torch.int64, device=self.device)
        sizes = [self.torch.zeros_like(n) for _ in range(self.world)]
        self.dist.all_gather(sizes, n, group=self.group)
        sizes = np.array([int(x.item()) for x in sizes], np.int64)
        m = int(sizes.max())
        view = arr.view(np.uint8).reshape(arr.shape[0], -1) if arr.shape[0] else np.zeros((0, arr.dtype.itemsize), np.uint8)
        pad = np.zeros((m, arr.dtype.itemsize), np.uint8)
        pad[:arr.shape[0]] = view
        t = self._to(pad)
        outs = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(outs, t, group=self.group)
        parts = [o.cpu().numpy()[:sizes[r]].reshape(-1).view(arr.dtype) for r, o in enumerate(outs)]
        return np.concatenate(parts) if parts else arr, sizes

    def allgather_object(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def allgather_fixed(self, arr):
        """all-gather of equally shaped int64 / float64 arrays -> array of shape (world,) + arr.shape: one collective,
        one copy to the device and one back (allgather_concat pays a size exchange and a copy per rank)"""
        arr = np.ascontiguousarray(arr)
        t = self._to(arr.reshape(-1).view(np.int64))
        out = self.torch.empty(self.world * t.numel(), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, t, group=self.group)
        return out.cpu().numpy().view(arr.dtype).reshape((self.world,) + arr.shape)

    def alltoall(self, arr, send_counts, recv_counts):
        """variable-size all-to-all of a 1-d numpy array: send_counts[r] consecutive elements go to rank r; returns the
        received elements in source-rank order (recv_counts[r] from rank r)"""
        arr = np.ascontiguousarray(arr)
        it = arr.dtype.itemsize
        src = self._to(arr.view(np.uint8).reshape(-1))
        out = self.torch.empty(int(np.sum(recv_counts)) * it, dtype=self.torch.uint8, device=self.device)
        self.dist.all_to_all_single(out, src, [int(c) * it for c in recv_counts], [int(c) * it for c in send_counts],
                                    group=self.group)
        return out.cpu().numpy().view(arr.dtype)

    def alltoall_dev(self, t, send_counts, recv_counts):
        """the same for a device tensor (first sum(send_counts) elements valid); stays on the device (NCCL)"""
        n_out = int(np.sum(recv_counts))
        out = self.torch.empty(max(1, n_out), dtype=t.dtype, device=t.device)
        self.dist.all_to_all_single(out[:n_out], t[:int(np.sum(send_counts))], [int(c) for c in recv_counts],
                                    [int(c) for c in send_counts], group=self.group)
        return out

    def allreduce_dev_inplace(self, t):
        """in-place sum over ranks of a device tensor (NCCL over NVLink), no host copy"""
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)

    def wrap_dev(self, ptr: int, n: int):
        """int64 tensor view of n elements of raw device memory (the walk's count vector), zero copy"""
        class _Raw:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}
        return self.torch.as_tensor(_Raw(), device=self.device)

    def allreduce_dev(self, t):
        """in-place sum over ranks of a device tensor (NCCL), returned as a numpy array"""
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    @property
    def on_gpu(self) -> bool:
        return self.device.type == "cuda"

    @property
    def peer_ok(self) -> bool:
        """one GPU per rank on one box (at most 8): the walk may use peer memory (mpb_peer_*) instead of a collective
        call per round.  Ranks that share a GPU (gloo) would spin on each other across time slices."""
        return self.device.type == "cuda" and self.world <= 8

    def empty_dev(self, n: int, dtype):
        """uninitialised device tensor of a numpy dtype (haplotype entries stay on the GPU between export and merge)"""
        tdt = {np.dtype(np.uint64): self.torch.int64, np.dtype(np.uint32): self.torch.int32,
               np.dtype(np.int64): self.torch.int64}[np.dtype(dtype)]
        return self.torch.empty(max(1, n), dtype=tdt, device=self.device)

    def allgather_dev(self, t, n: int, max_n: int):
        """all-gather of device tensors of different lengths (first n elements valid) -> list of per-rank tensors,
        each padded to max_n"""
        if t.shape[0] < max_n:
            pad = self.torch.empty(max_n, dtype=t.dtype, device=t.device)
            pad[:n] = t[:n]
            t = pad
        out = self.torch.empty(self.world * max_n, dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, t[:max_n].contiguous(), group=self.group)
        return [out[r * max_n:(r + 1) * max_n] for r in range(self.world)]

    def barrier(self):
        self.dist.barrier(group=self.group)
