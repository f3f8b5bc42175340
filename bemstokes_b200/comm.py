"""Exchange steps of the multi-GPU solve as C callbacks over torch.distributed (NCCL over NVLink on the GPU
box, gloo in the CPU tests).  The reference gets these from Epetra: Import in vmult = allgather of the Krylov
vector, Allreduce in Dot/Norm.  PyTorch is plumbing here: device pointers are wrapped without
copies and the collectives are ordered on the context's stream."""
import contextlib
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


class _DevPtr:
    """Zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def dev_tensor(ptr, n, device):
    """float64 tensor aliasing `n` doubles at `ptr` (device memory on a CUDA device, host memory on 'cpu')."""
    if device is None or torch.device(device).type == "cpu":
        buf = (C.c_double * n).from_address(int(ptr))
        return torch.from_numpy(np.ctypeslib.as_array(buf))
    return torch.as_tensor(_DevPtr(ptr, n), device=device)


class TorchComm:
    """Owns the two ctypes callbacks handed to bs_set_comm; keep the object alive as long as the context."""

    def __init__(self, group=None, device=None):
        self.group = group
        self.device = device
        self.n_allgather = 0
        self.n_allreduce = 0
        self._scratch = None
        self.allgatherv_cb = _lib.ALLGATHERV_FN(self._allgatherv)
        self.allreduce_cb = _lib.ALLREDUCE_FN(self._allreduce)

    def attach(self, ctx):
        _lib.check(_lib.lib.bs_set_comm(ctx, self.allgatherv_cb, self.allreduce_cb, None))

    def _on_stream(self, stream):
        """Order torch's work on the stream the library launches on (the `stream` argument of the callback): the
        collectives read and write the same device buffers as the library's kernels."""
        if stream and self.device is not None and torch.device(self.device).type == "cuda":
            return torch.cuda.stream(torch.cuda.ExternalStream(int(stream), device=self.device))
        return contextlib.nullcontext()

    def _allgatherv(self, user, send, sendcount, recv, counts, displs, stream):
        try:
            with self._on_stream(stream):
                P = dist.get_world_size(self.group)
                cnt = [counts[r] for r in range(P)]
                dsp = [displs[r] for r in range(P)]
                mx = max(cnt)
                total = max(d + c for d, c in zip(dsp, cnt))
                if self._scratch is None or self._scratch.numel() < (P + 1) * mx:
                    self._scratch = torch.zeros((P + 1) * mx, dtype=torch.float64, device=self.device)
                pad = self._scratch[P * mx:(P + 1) * mx]
                pad[:sendcount].copy_(dev_tensor(send, sendcount, self.device))
                out = self._scratch[:P * mx]
                dist.all_gather_into_tensor(out, pad, group=self.group)
                dst = dev_tensor(recv, total, self.device)
                for r in range(P):
                    dst[dsp[r]:dsp[r] + cnt[r]].copy_(out[r * mx:r * mx + cnt[r]])
            self.n_allgather += 1
            return 0
        except Exception as e:  # never let an exception cross the C boundary
            print("allgatherv callback failed:", repr(e), flush=True)
            return 1

    def _allreduce(self, user, buf, count, stream):
        try:
            with self._on_stream(stream):
                t = dev_tensor(buf, count, self.device)
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self.n_allreduce += 1
            return 0
        except Exception as e:
            print("allreduce callback failed:", repr(e), flush=True)
            return 1


def partition_ranges(n_nodes, nranks):
    """The library's default row partition: contiguous, balanced ranges of its locality order
    (mirrors build_geometry in csrc/bs_host.cu; host logic testable without a GPU)."""
    return [(n_nodes * r) // nranks for r in range(nranks + 1)]
