"""The multi-GPU exchange inside libmscnn_b200.so: final detections packed on the device (header with the
per-image counts + compacted rows) and ONE ncclAllGather on the communicator's own stream.

* one rank (any box): mscnn_net_detect_gather's payload == mscnn_net_detect's padded output, bit for bit;
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
    r = subprocess.run([str(exe), str(proto), "--gpus", str(n), "--steps", "4", "--warmup", "2", "--verify"],
                       capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = r.stdout.splitlines()
    ranks = [l for l in lines if l.startswith("rank ")]
    assert len(ranks) == n and all("gathered payload verified" in l for l in ranks)
    assert all(int(l.split(" proposals")[0].split()[-1]) > 0 for l in ranks), "degenerate run: no proposals"
    assert any(l.startswith("images_per_s") for l in lines)
