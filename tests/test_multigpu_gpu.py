"""The multi-GPU exchange inside libmscnn_b200.so: final detections packed on the device (header with the
per-image counts + compacted rows) and ONE ncclAllGather on the communicator's own stream.

* one rank (any box): mscnn_net_detect_gather's (NCCL) and mscnn_net_detect_push's (peer memory) payloads ==
  mscnn_net_detect's padded output, bit for bit; two PROCESSES exchanging through CUDA-IPC-mapped buffers;
* examples/multi_gpu_driver.cpp: ONE C++ process, one host thread per GPU (the reference's thread-local Caffe context,
  /root/reference/src/caffe/common.cpp:13-22), every rank checks the gathered slots of all ranks (--verify).  Uses every
  visible GPU (2, 4, 8 under `gpurun --gpus N`; a single GPU still runs the NCCL path with one rank).
The world_size-2 host logic (payload layout, rank-major merge) is covered on CPU by tests/test_parallel_cpu.py."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_detect_gather_single_rank_equals_detect(cuda):
    import torch
    from mscnn_b200 import models, net as mnet, parallel, synth
    mnet.set_precision("fp32")
    mnet.set_stream(torch.cuda.current_stream().cuda_stream)
    b, h, w = 3, 96, 320
    net = mnet.Net(models.kitti(h, w, 7, False, batch=b))
    net.set_params(synth.make_weights(net.layers()))
    net.forward(data=synth.make_images(b, h, w))
    cfg = mnet.kitti_detect_cfg(h, w)
    cap = cfg.max_rois_per_image
    dets = torch.zeros((b, cap, 5), device=cuda)
    cnt = torch.zeros(b, dtype=torch.int32, device=cuda)
    net.detect(cfg, dets.data_ptr(), cnt.data_ptr())
    comm = parallel.Comm(rank=0, world=1)
    per = parallel.payload_floats(b, cap)
    payload = torch.zeros(per, device=cuda)
    for _ in range(3):           # back-to-back gathers: each producer waits for the previous collective
        net.detect_gather(cfg, comm, payload.data_ptr())
    comm.synchronize()
    torch.cuda.synchronize()
    got = parallel.unpack_payload(payload, 1, b, cap)
    dets, cnt = dets.cpu().numpy(), cnt.cpu().numpy()
    assert sum(len(g) for g in got) == int(cnt.sum()) > 0
    for i in range(b):
        assert np.array_equal(got[i], dets[i, :cnt[i]]), i
    comm.close()


def test_detect_push_single_rank_equals_detect(cuda):
    """Peer-memory exchange with one rank: the fused post-process + push kernel writes the packed payload into the
    rank's own gather buffer (both generations over consecutive steps) == mscnn_net_detect's padded output."""
    import torch
    from mscnn_b200 import models, net as mnet, parallel, synth
    mnet.set_precision("fp32")
    mnet.set_stream(torch.cuda.current_stream().cuda_stream)
    b, h, w = 3, 96, 320
    net = mnet.Net(models.kitti(h, w, 7, False, batch=b))
    net.set_params(synth.make_weights(net.layers()))
    cfg = mnet.kitti_detect_cfg(h, w)
    cap = cfg.max_rois_per_image
    x = parallel.PeerExchange(b, cap, rank=0, world=1, generations=2)
    for step in range(5):
        net.forward(data=synth.make_images(b, h, w, first_index=step * b))
        dets = torch.zeros((b, cap, 5), device=cuda)
        cnt = torch.zeros(b, dtype=torch.int32, device=cuda)
        net.detect(cfg, dets.data_ptr(), cnt.data_ptr())
        net.detect_push(cfg, x)
        x.wait(torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got = parallel.unpack_payload(x.gathered(), 1, b, cap)
        d, c = dets.cpu().numpy(), cnt.cpu().numpy()
        assert int(c.sum()) > 0
        for i in range(b):
            assert np.array_equal(got[i], d[i, :c[i]]), (step, i)
    x.close()


def _ipc_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mscnn_b200 import models, net as mnet, parallel, synth
        torch.cuda.set_device(0)                     # both ranks share the one GPU: the buffers still cross PROCESSES
        mnet.set_device(0)
        mnet.set_precision("fp32")
        b, h, w = 2, 96, 320
        net = mnet.Net(models.kitti(h, w, 7, False, batch=b))
        net.set_params(synth.make_weights(net.layers()))
        cfg = mnet.kitti_detect_cfg(h, w)
        cap = cfg.max_rois_per_image
        x = parallel.PeerExchange(b, cap, generations=3)
        ok = True
        for step in range(7):
            first = (step * world + rank) * b
            net.forward(data=synth.make_images(b, h, w, first_index=first))
            dets = torch.zeros((b, cap, 5), device="cuda")
            cnt = torch.zeros(b, dtype=torch.int32, device="cuda")
            net.detect(cfg, dets.data_ptr(), cnt.data_ptr())
            net.detect_push(cfg, x)
            x.wait(0)
            torch.cuda.synchronize()
            mine = [dets[i, :int(cnt[i])].cpu().numpy() for i in range(b)]
            everyone = [None] * world
            dist.all_gather_object(everyone, mine)
            got = parallel.unpack_payload(x.gathered(), world, b, cap)
            for r in range(world):
                for i in range(b):
                    ok = ok and np.array_equal(got[r * b + i], everyone[r][i]) and len(everyone[r][i]) > 0
            dist.barrier()                            # consume before anybody's next push (generation reuse)
        x.close()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_detect_push_two_processes_over_cuda_ipc(cuda):
    """The multi-process path (one process per GPU under torchrun) on whatever this box has: two PROCESSES exchange
    their payloads through CUDA-IPC-mapped buffers (here on one device; over NVLink with one GPU each), seven steps over
    three generations, so every generation is reused; each rank checks every rank's slot against what that rank computed."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ipc_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=120)
    assert res == [(0, True), (1, True)]


def test_cpp_multi_gpu_driver_one_thread_per_gpu(cuda, tmp_path):
    import torch
    from mscnn_b200 import models
    exe = tmp_path / "multi_gpu_driver"
    cmd = ["g++", "-std=c++17", "-O2", "-I", str(ROOT / "include"), "-I", "/usr/local/cuda/include",
           str(ROOT / "examples/multi_gpu_driver.cpp"), "-L", str(ROOT / "mscnn_b200"), "-lmscnn_b200",
           "-L", "/usr/local/cuda/lib64", "-lcudart", "-lpthread", f"-Wl,-rpath,{ROOT / 'mscnn_b200'}", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    proto = tmp_path / "deploy.prototxt"
    proto.write_text(models.kitti(192, 640, 8, False, batch=2))
    n = torch.cuda.device_count()
    for exchange in ("peer", "nccl"):
        r = subprocess.run([str(exe), str(proto), "--gpus", str(n), "--steps", "4", "--warmup", "2", "--verify",
                            "--exchange", exchange], capture_output=True, text=True, timeout=600)
        print(r.stdout)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        lines = r.stdout.splitlines()
        ranks = [l for l in lines if l.startswith("rank ")]
        assert len(ranks) == n and all("gathered payload verified" in l for l in ranks), exchange
        assert all(int(l.split(" proposals")[0].split()[-1]) > 0 for l in ranks), "degenerate run: no proposals"
        assert any(l.startswith("images_per_s") for l in lines)
