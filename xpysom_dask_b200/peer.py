"""One-shot all-reduce of the per-epoch ``[S | c]`` partials over NVLink peer memory.

The sharded path has exactly one exchange step per epoch: the sum of the per-rank partial updates (the
reference does it with a Dask ``sum`` over its per-block ``_update`` results, xpysom.py:574-583).  For the
small buffers of most maps (config 2: 0.26 MB) an NCCL all-reduce is pure latency, so on a single node
the ranks open each other's mailboxes through CUDA IPC once (handles travel through the process group)
and ``libsom_b200``'s ``som_b200_peer_allreduce`` does the sum in one kernel on the compute stream.
Large buffers, several nodes, CUDA-graph replay or any set-up failure on ANY rank: NCCL is used instead
(the decision is agreed on by an all-reduce, so the ranks never disagree).

Measured (tools/peer_check.py, B200, back-to-back calls): 2 GPUs 16 us vs NCCL 20 us at 0.27 MB; 8 GPUs 40 us vs
NCCL 28 us (NVLS reduces inside the switch; this kernel reads seven remote mailboxes one after the other).  NCCL
therefore stays the default and this path is opt-in (``SOM_B200_PEER=1``).
"""
import ctypes
import socket
import zlib

import torch

from . import _lib

# above this size the 'every rank reads every mailbox' pattern loses to NCCL's ring / NVLS all-reduce (measured on
# 2 B200s: 16 us vs 20 us at 0.27 MB, 50 us vs 37 us at 4 MB): config 4's 31 MB buffer stays on NCCL
ONE_SHOT_MAX_BYTES = 1 << 20


class PeerReducer:
    """In-place sum of a device fp32 buffer across the ranks of ``group`` (all on one node)."""

    def __init__(self, eng, group, floats):
        import torch.distributed as dist
        self.eng, self.group, self.floats = eng, group, int(floats)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.comm = None
        ok = 1
        handle = (ctypes.c_uint8 * 64)()
        comm = ctypes.c_void_p()
        try:
            if self.world > 16:
                raise _lib.SomB200Error("peer all-reduce disabled")
            with torch.cuda.device(eng.device):
                _lib.check(eng.lib.som_b200_peer_create(self.floats, self.world, self.rank, ctypes.byref(comm),
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

    def all_reduce_(self, t):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() <= self.floats
        eng = self.eng
        with torch.cuda.device(eng.device):
            _lib.check(eng.lib.som_b200_peer_allreduce(self.comm, ctypes.c_void_p(t.data_ptr()), t.numel(),
                                                       eng._stream()), "som_b200_peer_allreduce")
        eng.launches += 1

    def close(self):
        if self.comm is not None:
            self.eng.lib.som_b200_peer_destroy(self.comm)
            self.comm = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
