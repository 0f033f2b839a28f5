"""Multi-GPU check of the one-shot NVLink all-reduce (run under torchrun on >= 2 GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/peer_check.py

Compares it with NCCL on several sizes (values, bitwise agreement across ranks, latency) and trains a small
map through XPySom both ways.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xpysom_dask_b200 import XPySom            # noqa: E402
from xpysom_dask_b200.engine import CudaEngine  # noqa: E402
from xpysom_dask_b200.peer import PeerReducer   # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
dev = torch.device("cuda", torch.cuda.current_device())
eng = CudaEngine(dev)
ok = True


def log(*a):
    if rank == 0:
        print(*a, flush=True)


for n in (1, 5, 27200, 1024 * 65, 1 << 20):
    red = PeerReducer(eng, dist.group.WORLD, n)
    if not red.active:
        log("peer all-reduce NOT active (IPC unavailable?) -> NCCL fallback would be used")
        ok = False
        break
    g = torch.Generator(device="cpu").manual_seed(1000 * n + rank)
    worst = 0.0
    for it in range(20):
        x = torch.randn(n, generator=g).to(dev)
        ref = x.clone()
        dist.all_reduce(ref)
        red.all_reduce_(x)
        worst = max(worst, float((x - ref).abs().max() / ref.abs().max().clamp_min(1e-30)))
        gathered = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(gathered, x)
        same = all(torch.equal(gathered[0], t) for t in gathered)
        if not same or worst > 1e-5:
            ok = False
    # latency, back to back on one stream
    x = torch.randn(n, generator=g).to(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    reps = 200
    torch.cuda.synchronize(); dist.barrier()
    ev[0].record()
    for _ in range(reps):
        red.all_reduce_(x)
    ev[1].record()
    torch.cuda.synchronize(); dist.barrier()
    ev[2].record()
    for _ in range(reps):
        dist.all_reduce(x)
    ev[3].record()
    torch.cuda.synchronize()
    log("n=%8d floats: max rel diff vs NCCL %.2e, identical on all ranks %s | one-shot %.1f us, NCCL %.1f us per call"
        % (n, worst, same, ev[0].elapsed_time(ev[1]) * 1e3 / reps, ev[2].elapsed_time(ev[3]) * 1e3 / reps))
    red.close()

# the host class, both ways
# deterministic set-up (small-integer samples: exact per-BMU sums whatever the order of the atomics; a 64-neuron map:
# un-sliced neighbourhood apply), so that two runs can be compared tightly instead of drifting apart chaotically
rng = np.random.RandomState(21)
centres = rng.randint(0, 8, size=(24, 32))
data = (centres[rng.randint(24, size=16000)] + rng.randint(0, 2, size=(16000, 32))).astype(np.float32)
shard = data[rank::world]
res = {}
for mode in ("peer", "nccl"):
    os.environ["SOM_B200_PEER"] = "0" if mode == "nccl" else "1"
    som = XPySom(8, 8, 32, sigma=2.0, random_seed=4, device=dev, process_group=True)
    som.train(shard, 6)
    res[mode] = torch.as_tensor(som.get_weights()).to(dev)
    used = getattr(som, "_peer_cache", None)
    log("XPySom.train (%s): peer reducer %s" % (mode, "active" if used is not None and used[1].active else "not used"))
diff = float((res["peer"] - res["nccl"]).abs().max() / res["nccl"].abs().max())
gath = [torch.empty_like(res["peer"]) for _ in range(world)]
dist.all_gather(gath, res["peer"])
replicated = all(torch.equal(gath[0], t) for t in gath)
log("codebook after 6 epochs: peer vs NCCL max rel diff %.2e; replicated bit-exactly across ranks: %s" % (diff, replicated))
if diff > 1e-6 or not replicated:
    ok = False
log("PEER CHECK", "PASSED" if ok else "FAILED")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
