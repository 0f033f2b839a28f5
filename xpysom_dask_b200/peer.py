"""Exact accumulators in NVLink peer memory: the sharded path's one exchange step, fused into the epoch tail.

The sharded path has exactly one exchange step per epoch: the sum of the per-rank accumulators (the reference does
it with a Dask ``sum`` over its per-block ``_update`` results, xpysom.py:545-558).  For the small buffers of most
maps (config 2: 0.5 MB) an NCCL all-reduce is pure latency and a stream hand-over in the middle of the epoch's launch
chain, so on a single node the ranks open each other's accumulators through CUDA IPC once (handles travel through the
process group) and the finalize phase of ``som_b200_epoch_tail`` sums them over the ranks itself (csrc/peer.cuh).
Large buffers, several nodes, CUDA-graph replay or any set-up failure on ANY rank: NCCL all-reduces the integer
accumulator instead (the decision is agreed on by an all-reduce, so the ranks never disagree).
"""
import ctypes
import socket
import zlib

import torch

from . import _lib

# above this size the 'every rank reads every accumulator' pattern loses to NCCL's ring / NVLS all-reduce
# (world - 1 remote reads of the whole buffer per rank): config 4's 63 MB accumulator stays on NCCL
PEER_MAX_BYTES = 4 << 20


class PeerAccumulator:
    """Accumulators of ``words`` 64-bit words on every rank of ``group`` (all on one node), readable by all of them."""

    def __init__(self, eng, group, words):
        import torch.distributed as dist
        self.eng, self.group, self.words = eng, group, int(words)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.comm = None
        ok = 1
        handle = (ctypes.c_uint8 * 64)()
        comm = ctypes.c_void_p()
        try:
            if self.world > 16 or self.world < 2:
                raise _lib.SomB200Error("peer accumulators disabled")
            with torch.cuda.device(eng.device):
                _lib.check(eng.lib.som_b200_peer_create(self.words, self.world, self.rank, ctypes.byref(comm),
                                                        ctypes.cast(handle, ctypes.c_void_p)), "som_b200_peer_create")
        except _lib.SomB200Error:
            ok = 0
        # every rank sends (ok, host id, handle); the mailboxes are only usable when all ranks share a host
        host = zlib.crc32(socket.gethostname().encode())        # same value on every rank of one host
        mine = torch.tensor([ok, host] + list(bytes(handle)), dtype=torch.int64, device=eng.device)
        allv = torch.empty(self.world * mine.numel(), dtype=torch.int64, device=eng.device)
        dist.all_gather_into_tensor(allv, mine, group=group)
        allv = allv.view(self.world, -1).cpu()
        usable = bool((allv[:, 0] == 1).all()) and bool((allv[:, 1] == allv[0, 1]).all())
        if usable:
            blob = bytes(allv[:, 2:].to(torch.uint8).flatten().tolist())
            try:
                with torch.cuda.device(eng.device):
                    _lib.check(eng.lib.som_b200_peer_connect(comm, ctypes.c_char_p(blob)), "som_b200_peer_connect")
            except _lib.SomB200Error:
                usable = False
        # second agreement: a rank whose cudaIpcOpenMemHandle failed must take everyone to NCCL with it
        flag = torch.tensor([1 if usable else 0], dtype=torch.int32, device=eng.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 1:
            self.comm = comm
        elif comm.value:
            eng.lib.som_b200_peer_destroy(comm)

    @property
    def active(self):
        return self.comm is not None

    def current(self):
        """Device address of the accumulator of the NEXT epoch (alternates: ask every epoch)."""
        p = self.eng.lib.som_b200_peer_accumulator(self.comm)
        if not p:
            raise _lib.SomB200Error("som_b200_peer_accumulator: communicator not connected")
        return ctypes.c_void_p(p)

    def fence(self):
        """Every rank has finished reading every accumulator (call before the ranks may diverge: end of train())."""
        import torch.distributed as dist
        t = torch.zeros(1, dtype=torch.int32, device=self.eng.device)
        dist.all_reduce(t, group=self.group)

    def close(self):
        """Collective: all ranks close together (no peer may still be inside an exchange)."""
        if self.comm is not None:
            try:
                self.fence()
                torch.cuda.synchronize(self.eng.device)
            finally:
                self.eng.lib.som_b200_peer_destroy(self.comm)
                self.comm = None

    def __del__(self):
        # every train() call ends with fence(): nobody is inside an exchange when the object is collected
        try:
            if self.comm is not None:
                self.eng.lib.som_b200_peer_destroy(self.comm)
                self.comm = None
        except Exception:
            pass
