"""world_size-2 gloo test of the multi-GPU plumbing (sharding + the single all-gather of final
detections) on CPU tensors; the NCCL path runs the same functions on the GPU box (bench.py --gpus N)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mscnn_b200 import parallel
    b, cap = 3, 16
    first, last = parallel.shard_range(rank, b)
    dets = torch.zeros((b, cap, 5))
    counts = torch.zeros(b, dtype=torch.int32)
    for i, g in enumerate(range(first, last)):      # image g has g+1 detections whose x encodes (g, k)
        counts[i] = g + 1
        for k in range(g + 1):
            dets[i, k] = torch.tensor([100.0 * g + k, 1, 2, 3, 0.5])
    buf = parallel.GatherBuffers(world, b, cap, "cpu")
    parallel.all_gather_detections(dets, counts, buf)
    merged = parallel.merge_detections(buf)
    ok = len(merged) == world * b and all(
        m.shape == (g + 1, 5) and [float(v) for v in m[:, 0]] == [100.0 * g + k for k in range(g + 1)]
        for g, m in enumerate(merged))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_shard_range():
    from mscnn_b200 import parallel
    assert parallel.shard_range(0, 8) == (0, 8) and parallel.shard_range(3, 8) == (24, 32)
