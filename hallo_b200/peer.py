"""Peer-mapped exchange memory of a frame-sharded window (one process per GPU, NVLink / NVSwitch).

Each rank owns one `hallo_b200_peer_alloc` allocation (cudaMalloc, exportable through CUDA IPC), carved into the same
named regions at the same offsets on every rank, so "region X on rank r" is `peer_base[r] + offset(X)`.  The kernels
that produce data for another rank store it there directly (GroupNorm-apply scatter, GEMM row-scatter epilogue:
include/hallo_b200.h "Peer memory"); `barrier()` is the flag barrier that orders those stores against the consumers.
torch.distributed is used once, to exchange the 64-byte IPC handles.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import torch

from . import lib


class _DevMem:
    """Minimal __cuda_array_interface__ carrier: lets torch view library-owned device memory without copying."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class PeerArena:
    FLAG_BYTES = 256            # HB_MAX_PEERS uint32 flag words, padded

    def __init__(self, regions: List[Tuple[str, int]], group, rank: int, world: int, device):
        """regions: [(name, bytes)] -- identical on every rank (same shapes everywhere)."""
        import torch.distributed as dist
        self.rank, self.world, self.device = rank, world, device
        self.offsets: Dict[str, Tuple[int, int]] = {}
        off = self.FLAG_BYTES
        for name, nbytes in regions:
            off = (off + 1023) // 1024 * 1024
            self.offsets[name] = (off, nbytes)
            off += nbytes
        self.nbytes = (off + 1023) // 1024 * 1024
        h = lib.load()
        # Every rank runs the SAME sequence of collectives whatever fails locally; the verdict is agreed with a MIN
        # all-reduce, so either all ranks own a fully mapped arena or all of them raise.
        self.local_base, self.bases, self._opened, self._bytes = 0, [], [], None
        err = None
        mine = b""
        try:
            base = C.c_void_p()
            lib.check(h.hallo_b200_peer_alloc(C.c_int64(self.nbytes), C.byref(base)), "peer_alloc")
            self.local_base = int(base.value)
            handle = (C.c_ubyte * 64)()
            lib.check(h.hallo_b200_peer_export(C.c_void_p(self.local_base), handle), "peer_export")
            mine = bytes(handle)
        except Exception as e:
            err = e
        handles: List[bytes] = [b""] * world
        dist.all_gather_object(handles, mine, group=group)
        if err is None and all(len(x) == 64 for x in handles):
            try:
                for r in range(world):
                    if r == rank:
                        self.bases.append(self.local_base)
                        continue
                    buf = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                    p = C.c_void_p()
                    lib.check(h.hallo_b200_peer_open(buf, C.byref(p)), f"peer_open(rank {r})")
                    self.bases.append(int(p.value))
                    self._opened.append(int(p.value))
            except Exception as e:
                err = e
        elif err is None:
            err = RuntimeError("another rank could not export its exchange buffer")
        ok = torch.tensor([0 if err is not None else 1], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok) == 0:
            self.close()
            raise RuntimeError(f"peer-memory arena unavailable: {err if err is not None else 'failed on another rank'}")
        self._mem = _DevMem(self.local_base, self.nbytes)
        self._bytes = torch.as_tensor(self._mem, device=device)           # uint8 view of the local allocation
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self._flag_ptrs = (C.c_void_p * world)(*[C.c_void_p(b) for b in self.bases])
        dist.barrier(group=group)                                          # every rank has mapped every allocation

    # ------------------------------------------------------------------ views / addresses
    def local(self, name: str, shape, dtype) -> torch.Tensor:
        off, nbytes = self.offsets[name]
        n = 1
        for s in shape:
            n *= s
        esz = torch.empty(0, dtype=dtype).element_size()
        assert n * esz <= nbytes, (name, shape, nbytes)
        return self._bytes[off:off + n * esz].view(dtype).view(*shape)

    def addr(self, name: str, r: int, byte_offset: int = 0) -> int:
        return self.bases[r] + self.offsets[name][0] + byte_offset

    def addrs(self, name: str, byte_offset: int = 0) -> List[int]:
        return [self.addr(name, r, byte_offset) for r in range(self.world)]

    # ------------------------------------------------------------------ synchronisation
    def barrier(self):
        """Flag barrier on the current stream (graph-capturable): orders this rank's peer stores before it against
        every rank's kernels after it."""
        lib.check(lib.load().hallo_b200_peer_barrier(self._flag_ptrs, C.c_int(self.world), C.c_int(self.rank),
                                                     C.c_void_p(self.epoch.data_ptr()), lib.current_stream()),
                  "peer_barrier")

    def close(self):
        h = lib.load()
        torch.cuda.synchronize()
        self._bytes = None
        for p in self._opened:
            h.hallo_b200_peer_close(C.c_void_p(p))
        self._opened = []
        if self.local_base:
            h.hallo_b200_peer_free(C.c_void_p(self.local_base))
            self.local_base = 0
