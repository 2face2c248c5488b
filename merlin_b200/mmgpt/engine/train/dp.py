"""Data-parallel gradient exchange: bucketed all-reduce over slices of the flat gradient arena.

north_star: "the interpair / interleave multi-image batch shards over the 8 GPUs of one box with a single
NCCL allreduce on the gradient buckets (data-parallel only, no tensor parallel)".  Replaces the FSDP
reduce-scatter / all-gather of the reference (pretrain.sh:32-33).  Buckets are contiguous arena
ranges in backward-completion order; `ready(upto)` enqueues the all-reduce of every bucket that lies entirely
below `upto` on a side stream, so the exchange overlaps the remaining backward kernels; `finish()` makes the
compute stream wait for all of them.  Gradients are pre-scaled by 1/world_size in the loss backward, so SUM
yields the mean.

On CUDA the collective is issued by the library (`merlin_dp_init / merlin_dp_allreduce_bucket /
merlin_dp_finalize`, csrc/dp.cu): its own NCCL communicator with a CTA budget (`MERLIN_DP_CTAS`, default 8 with a
registered gradient arena (8 ranks and more), else 16),
and while buckets are in flight the persistent tcgen05 GEMMs are told to leave `MERLIN_DP_SM_RESERVE` SMs free
(`merlin_set_sm_budget`) — a persistent GEMM that claims all 148 SMs cannot co-reside with NCCL's CTAs, its
static tile schedule then waits for the SMs the collective holds and the exchange is serialised with backward
instead of hidden under it (round-1 finding: ~100 % of the all-reduce time was exposed).  torch.distributed is
used only as plumbing: to hand the 128-byte NCCL unique id from rank 0 to the other ranks.
On CPU tensors (the gloo tests) the same bucket logic runs over torch.distributed collectives.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def _torch_nccl_path() -> Optional[str]:
    """The libnccl.so.2 PyTorch itself links against (so the library and torch share one NCCL)."""
    here = os.path.dirname(torch.__file__)
    for cand in (os.path.join(here, "..", "nvidia", "nccl", "lib", "libnccl.so.2"),
                 os.path.join(here, "lib", "libnccl.so.2")):
        if os.path.exists(cand):
            return os.path.abspath(cand)
    return None


class LibraryComm:
    """NCCL communicator owned by libmerlin_b200.so (merlin_dp_*)."""

    def __init__(self, group=None, max_ctas: int = 0):
        from .... import lib as _lib
        self._lib, self._check = _lib.load(), _lib.check
        L = self._lib
        path = _torch_nccl_path()
        self._check(L.merlin_dp_load_nccl(path.encode() if path else None), "merlin_dp_load_nccl")
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        uid = ctypes.create_string_buffer(128)
        if self.rank == 0:
            self._check(L.merlin_dp_unique_id(uid), "merlin_dp_unique_id")
        box = [bytes(uid.raw)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        handle = ctypes.c_void_p()
        self._check(L.merlin_dp_init(box[0], self.rank, self.world, int(max_ctas), ctypes.byref(handle)),
                    "merlin_dp_init")
        self.handle, self.max_ctas = handle, int(max_ctas)

    def allreduce_bucket(self, t: torch.Tensor, stream: torch.cuda.Stream):
        assert t.dtype == torch.bfloat16 and t.is_contiguous()
        self._check(self._lib.merlin_dp_allreduce_bucket(self.handle, stream.cuda_stream, t.data_ptr(), t.numel()),
                    "merlin_dp_allreduce_bucket")

    def allreduce_f32(self, t: torch.Tensor, stream: torch.cuda.Stream, op_max: bool = False):
        assert t.dtype == torch.float32 and t.is_contiguous()
        self._check(self._lib.merlin_dp_allreduce_f32(self.handle, stream.cuda_stream, t.data_ptr(), t.numel(),
                                                      int(op_max)), "merlin_dp_allreduce_f32")

    def close(self):
        if self.handle:
            self._lib.merlin_dp_finalize(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover - interpreter shutdown order
        try:
            self.close()
        except Exception:
            pass


class NcclBuffer:
    """Device memory from ncclMemAlloc (merlin_dp_mem_alloc) exposed to torch through __cuda_array_interface__.
    NCCL can register such a buffer (merlin_dp_register): with NVLS the switch then reduces straight out of the
    gradient arena — no staging copies through NCCL's own buffers, i.e. less SM work per byte."""

    def __init__(self, nbytes: int):
        from .... import lib as _lib
        self._lib, self._check = _lib.load(), _lib.check
        ptr = ctypes.c_void_p()
        self._check(self._lib.merlin_dp_mem_alloc(ctypes.byref(ptr), int(nbytes)), "merlin_dp_mem_alloc")
        self.ptr, self.nbytes = ptr.value, int(nbytes)
        self.__cuda_array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False),
                                         "version": 3, "strides": None}

    def tensor(self, dtype: torch.dtype) -> torch.Tensor:
        return torch.as_tensor(self, device=torch.device("cuda", torch.cuda.current_device())).view(dtype)

    def free(self):
        if self.ptr:
            self._lib.merlin_dp_mem_free(ctypes.c_void_p(self.ptr))
            self.ptr = 0


class GradBucketReducer:
    def __init__(self, flat_grads: torch.Tensor, bucket_elems: int = 128 * 1024 * 1024, group=None,
                 backend: Optional[str] = None, max_ctas: Optional[int] = None, sm_reserve: Optional[int] = None):
        self.flat = flat_grads
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        n = flat_grads.numel()
        self.buckets: List[Tuple[int, int]] = [(a, min(n, a + bucket_elems)) for a in range(0, n, bucket_elems)]
        self.next_bucket = 0
        self.handles = []
        self.comm_stream = torch.cuda.Stream(priority=-1) if flat_grads.is_cuda else None
        self.bytes_reduced = 0
        self.exchange_enabled = True        # False: skip the collectives (compute-only timing, scripts/train_timeline.py)
        self.lib_comm: Optional[LibraryComm] = None
        self._budget_on = False
        backend = backend or os.environ.get("MERLIN_DP_BACKEND", "library")
        if backend not in ("library", "torch"):
            raise ValueError(f"MERLIN_DP_BACKEND={backend!r}: expected 'library' or 'torch'")
        # measured (profiles/r2_dp_overlap_sweep_n{2,4,8}.md).  8 ranks: NCCL runs NVLS; with the gradient arena
        # REGISTERED (ncclMemAlloc + ncclCommRegister: the switch reduces in place) 8 CTAs move the 13.55 GB under
        # backward (202 ms/step, 6 ms exposed; unregistered needs 16 CTAs and exposes 12-15 ms).  2 and 4 ranks: NCCL
        # picks the ring over P2P (~13 GB/s per CTA, registration changes nothing), so 16 CTAs: 8 CTAs expose 51 ms at
        # N = 4, 16 CTAs 11.5 ms.  Fewer CTAs than that are slower than backward itself.
        self.register = os.environ.get("MERLIN_DP_REGISTER", "1" if self.world >= 8 else "0") == "1"
        default_ctas = "8" if self.register else "16"
        self.max_ctas = int(os.environ.get("MERLIN_DP_CTAS", default_ctas)) if max_ctas is None else int(max_ctas)
        if sm_reserve is None:
            sm_reserve = int(os.environ.get("MERLIN_DP_SM_RESERVE", str((self.max_ctas + 1) // 2 * 2)))
        self.sm_reserve = int(sm_reserve)
        self.backend = "torch"
        self.registered = None
        if self.world > 1 and flat_grads.is_cuda and backend == "library":
            self.lib_comm = LibraryComm(group, self.max_ctas)
            self.backend = "library"

    def adopt_registered_arena(self, arena) -> bool:
        """MERLIN_DP_REGISTER (default: on from 8 ranks, where NCCL runs NVLS): move the gradient arena into an ncclMemAlloc'ed buffer and register it with the
        communicator (user-buffer registration: NVLS reduces in place).  Returns False when not applicable."""
        if self.lib_comm is None or not self.register:
            return False
        buf = NcclBuffer(arena.grads.numel() * arena.grads.element_size())
        t = buf.tensor(arena.grads.dtype)
        t.zero_()
        arena.grads = t
        self.flat = t
        handle = ctypes.c_void_p()
        L = self.lib_comm
        L._check(L._lib.merlin_dp_register(L.handle, ctypes.c_void_p(buf.ptr), buf.nbytes, ctypes.byref(handle)),
                 "merlin_dp_register")
        self.registered = (buf, handle)
        return True

    def start_step(self):
        self.next_bucket = 0
        self.handles = []

    def ready(self, upto: int):
        """Gradients in arena range [0, upto) are final: reduce every complete bucket below it."""
        if self.world == 1 or not self.exchange_enabled:
            return
        while self.next_bucket < len(self.buckets) and self.buckets[self.next_bucket][1] <= upto:
            a, b = self.buckets[self.next_bucket]
            self._launch(self.flat[a:b])
            self.next_bucket += 1

    def _set_budget(self, on: bool):
        """Persistent kernels launched from now on leave `sm_reserve` SMs to the collective's CTAs."""
        if self.lib_comm is None or self.sm_reserve <= 0 or on == self._budget_on:
            return
        L = self.lib_comm._lib
        L.merlin_set_sm_budget(L.merlin_sm_count() - self.sm_reserve if on else 0)
        self._budget_on = on

    def _launch(self, t: torch.Tensor):
        self.bytes_reduced += t.numel() * t.element_size()
        if self.comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record()                                    # compute stream: this bucket's grads are written
            self.comm_stream.wait_event(ev)
            if self.lib_comm is not None:
                self._set_budget(True)
                self.lib_comm.allreduce_bucket(t, self.comm_stream)
            else:
                with torch.cuda.stream(self.comm_stream):
                    self.handles.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self.handles.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Flush the remaining buckets and make the current stream wait for every all-reduce."""
        if self.world == 1 or not self.exchange_enabled:
            return
        self.ready(self.flat.numel())
        for h in self.handles:
            h.wait()
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self._set_budget(False)
        self.handles = []

    def close(self):
        self._set_budget(False)
        if self.registered is not None and self.lib_comm is not None:
            buf, handle = self.registered
            self.lib_comm._lib.merlin_dp_deregister(self.lib_comm.handle, handle)
            self.registered = None
        if self.lib_comm is not None:
            self.lib_comm.close()
            self.lib_comm = None
