"""Row sharding over the GPUs of one node: the collectives hb_bayes() asks its caller for, on torch.distributed.

One process per GPU (torchrun); every rank holds its individuals' rows of y and X.  `Comm` wraps the default
process group (NCCL on GPUs, gloo in the CPU tests) into the three C callbacks of hb_bayes_args
(include/hibayes_b200.h): a sum-all-reduce of host doubles, a sum-all-reduce of a device int32 buffer (the Gram
band) and an all-gather of small byte strings (CUDA IPC handles).  torch.distributed is plumbing only: the dots of
the sweep itself are exchanged inside the sweep kernel with NVLink peer atomics.
"""
import ctypes as C

import numpy as np

from . import _lib


class _DevView:
    """Exposes raw device memory to torch through __cuda_array_interface__."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 3}


class Comm:
    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.on_gpu = dist.get_backend(group) == "nccl"
        self._cbs = None

    def _tensor(self, arr):
        t = self.torch.from_numpy(arr)
        return t.cuda() if self.on_gpu else t

    def allreduce_f64(self, arr):
        """In-place sum over ranks of a host float64 array."""
        t = self._tensor(arr)
        self.dist.all_reduce(t, group=self.group)
        if self.on_gpu:
            arr[...] = t.cpu().numpy()
        return arr

    def total_rows(self, n_local):
        return int(self.allreduce_f64(np.array([float(n_local)]))[0])

    def allgather_bytes(self, mine):
        t = self._tensor(np.frombuffer(mine, dtype=np.uint8).copy())
        outs = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(outs, t, group=self.group)
        return b"".join(bytes(o.cpu().numpy().tobytes()) for o in outs)

    def allreduce_i32_device(self, ptr, count, chunk=1 << 28):
        """In-place sum over ranks of `count` int32 values in device memory (NCCL only)."""
        for off in range(0, count, chunk):
            c = min(chunk, count - off)
            t = self.torch.as_tensor(_DevView(ptr + 4 * off, c, "<i4"), device="cuda")
            self.dist.all_reduce(t, group=self.group)
        self.torch.cuda.synchronize()

    def callbacks(self):
        if self._cbs is None:
            def ar64(ctx, buf, count):
                try:
                    self.allreduce_f64(np.ctypeslib.as_array(buf, shape=(count,)))
                    return 0
                except Exception as e:  # noqa: BLE001 -- must not unwind through C
                    print("hibayes_b200.sharded: all-reduce failed:", e)
                    return 1

            def ar32(ctx, ptr, count):
                try:
                    self.allreduce_i32_device(int(ptr), int(count))
                    return 0
                except Exception as e:  # noqa: BLE001
                    print("hibayes_b200.sharded: device all-reduce failed:", e)
                    return 1

            def agb(ctx, mine, allp, nbytes):
                try:
                    data = self.allgather_bytes(C.string_at(mine, nbytes))
                    C.memmove(allp, data, len(data))
                    return 0
                except Exception as e:  # noqa: BLE001
                    print("hibayes_b200.sharded: all-gather failed:", e)
                    return 1

            self._cbs = (_lib.ALLREDUCE_F64(ar64), _lib.ALLREDUCE_I32_DEV(ar32), _lib.ALLGATHER_BYTES(agb))
        return self._cbs


def shard_rows(n, rank, world, multiple=4):
    """Contiguous row block of `rank`: boundaries at multiples of `multiple` (the synthetic generator addresses
    genotypes four rows at a time)."""
    per = -(-n // world)
    per = -(-per // multiple) * multiple
    lo = min(n, rank * per)
    return lo, min(n, lo + per)
