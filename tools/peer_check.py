"""Multi-GPU check of the sharded path's exchange step (run under torchrun on >= 2 GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/peer_check.py

Trains several maps through XPySom with the accumulators in NVLink peer memory (the exchange fused into the epoch
tail, csrc/peer.cuh) and with the NCCL all-reduce of the integer accumulator, and checks that both give the SAME bits,
that every rank holds the same bits, and that they equal ONE GPU training on all the rows; then times both on the
config-2 shard.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xpysom_dask_b200 import XPySom            # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
dev = torch.device("cuda", torch.cuda.current_device())
ok = True
quick = "--quick" in sys.argv


def log(*a):
    if rank == 0:
        print(*a, flush=True)


def same_everywhere(t):
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool(torch.equal(lo, hi))


CASES = [
    # (gx, gy, d, rows per rank, kwargs): fused cooperative tail | separate launches (wide) | hexagonal | tiny D
    (16, 16, 64, 20000, {}),
    (24, 25, 100, 12000, {"neighborhood_function": "bubble"}),
    (12, 11, 24, 9000, {"topology": "hexagonal", "activation_distance": "cosine"}),
    (40, 40, 16, 30000, {"decay_function": "linear"}),
    (9, 9, 7, 5000, {"activation_distance": "manhattan"}),
]
for gx, gy, d, n, kw in CASES:
    rng = np.random.RandomState(100 + rank)
    shard = torch.from_numpy(rng.random_sample((n + 37 * rank, d)).astype(np.float32)).to(dev)    # ragged shards
    res = {}
    for mode in ("peer", "peer_hot", "nccl"):
        os.environ["SOM_B200_PEER"] = "0" if mode == "nccl" else "1"
        som = XPySom(gx, gy, d, random_seed=4, device=dev, process_group=True, **kw)
        som._hot_bmus = mode == "peer_hot"       # forced: accumulate into local replicas, fold into the peer accumulator
        som.train(shard, 7)
        som._hot_bmus = mode == "peer_hot"
        som.train(shard, 12, iter_beg=7, iter_end=12)            # a second call on the same communicator
        res[mode] = torch.as_tensor(som.get_weights()).to(dev)
        used = getattr(som, "_peer_cache", None)
        active = used is not None and used[1].active
        if mode != "nccl" and not active:
            log("  peer accumulators NOT active (IPC unavailable?)")
            ok = False
        if used is not None:
            used[1].close()
            som._peer_cache = None
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([shard.shape[0]], dtype=torch.int64, device=dev))
    parts = [torch.empty((int(s.item()), d), dtype=torch.float32, device=dev) for s in sizes]
    dist.all_gather(parts, shard)
    single = XPySom(gx, gy, d, random_seed=4, device=dev, **kw)
    single.train(torch.cat(parts), 7)
    single.train(torch.cat(parts), 12, iter_beg=7, iter_end=12)
    one = torch.as_tensor(single.get_weights()).to(dev)
    a = torch.equal(res["peer"], res["nccl"]) and torch.equal(res["peer_hot"], res["nccl"])
    b = same_everywhere(res["peer"])
    c = torch.equal(res["peer"], one)
    rel = float((res["peer"] - one).abs().max() / one.abs().max())
    log("%2dx%-2d d=%-3d %s: peer == peer via local replicas == NCCL bitwise %s | identical on all ranks %s | == one GPU on all rows %s (rel %.1e)"
        % (gx, gy, d, kw or "", a, b, c, rel))
    if not (a and b and c):
        ok = False

if not quick:
    # timing on the config-2 shard (1M x 64 per rank, 32 x 32)
    x = torch.rand((1_000_000, 64), device=dev)
    for mode in ("peer", "nccl", "peer", "nccl"):
        os.environ["SOM_B200_PEER"] = "0" if mode == "nccl" else "1"
        som = XPySom(32, 32, 64, random_seed=0, device=dev, process_group=True)
        som.train(x, 200, iter_beg=0, iter_end=5)
        ts = []
        for rep in range(5):
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            som.train(x, 200, iter_beg=5, iter_end=55)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 50], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t.item()))
        log("config 2 shard on %d GPUs, %-4s: %.4f ms per epoch (median of 5 x 50 epochs; all: %s)"
            % (world, mode, float(np.median(ts)), " ".join("%.4f" % v for v in ts)))
        used = getattr(som, "_peer_cache", None)
        if used is not None:
            used[1].close()
            som._peer_cache = None

log("PEER CHECK", "PASSED" if ok else "FAILED")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
