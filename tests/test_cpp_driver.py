"""The drop-in boundary from the C++ side: examples/caffe_driver.cpp is written against the Caffe API only
(caffe::Net / Blob / Layer / LayerRegistry, /root/reference/include/caffe/net.hpp:23-120, layer_factory.hpp:56-137),
compiled with plain g++ against the Caffe-API mirror headers and linked against libmscnn_b200.so.

CPU (no GPU needed): it compiles, links, builds the net from an unchanged deploy prototxt, lists the same layers as the
C facade, and a layer type registered by the HOST program (REGISTER_LAYER_CLASS) lands in the library's registry.
GPU: Net::Forward() from C++ gives the outputs the Python facade gives for the same deterministic parameters."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _build(tmp_path):
    from mscnn_b200 import capi
    capi.lib()                                      # raises if the library has not been built
    exe = tmp_path / "caffe_driver"
    cmd = ["g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), "-I", str(ROOT / "mscnn_b200/csrc/caffe_api"),
           "-I", str(ROOT / "mscnn_b200/csrc/proto_shared"), "-I", "/usr/local/cuda/include",
           str(ROOT / "examples/caffe_driver.cpp"), "-L", str(ROOT / "mscnn_b200"), "-lmscnn_b200",
           f"-Wl,-rpath,{ROOT / 'mscnn_b200'}", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def _fill(count, salt, scale):
    e = np.arange(count, dtype=np.uint64)
    h = ((e * np.uint64(2654435761) + np.uint64(salt) * np.uint64(40503)) & np.uint64(0xFFFFFFFF)) >> np.uint64(8)
    v = (h & np.uint64(0xFFFF)).astype(np.float32) / np.float32(65536.0) - np.float32(0.5)
    return v * np.float32(scale)


def test_cpp_host_compiles_links_and_builds_the_net(tmp_path):
    from mscnn_b200 import models, net as mnet
    exe = _build(tmp_path)
    proto = tmp_path / "deploy.prototxt"
    proto.write_text(models.kitti(96, 320, 7, False, batch=1))
    r = subprocess.run([str(exe), str(proto)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.splitlines()
    net = mnet.Net(str(proto))
    got = [(l.split()[2], l.split()[3]) for l in lines if l.startswith("layer ")]
    assert got == list(zip(net.layer_names, net.layer_types))
    assert any(l.startswith("registry:") and "HostPass=1 BoxOutput=1" in l for l in lines)


def test_multi_gpu_driver_compiles_and_links_without_a_gpu(tmp_path):
    """examples/multi_gpu_driver.cpp (one host thread per GPU, C ABI only: net facade, peer-memory exchange, NCCL
    communicator) builds with plain g++ against include/mscnn_b200.h and links against the library; without a device
    it stops with its usage / device-count message instead of crashing."""
    from mscnn_b200 import capi
    capi.lib()
    exe = tmp_path / "multi_gpu_driver"
    cmd = ["g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), "-I", "/usr/local/cuda/include",
           str(ROOT / "examples/multi_gpu_driver.cpp"), "-L", str(ROOT / "mscnn_b200"), "-lmscnn_b200",
           "-L", "/usr/local/cuda/lib64", "-lcudart", "-lpthread", f"-Wl,-rpath,{ROOT / 'mscnn_b200'}", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "usage:" in r.stderr


@pytest.mark.gpu
def test_cpp_host_forward_matches_the_facade(cuda, tmp_path):
    from mscnn_b200 import models, net as mnet
    exe = _build(tmp_path)
    proto = tmp_path / "deploy.prototxt"
    proto.write_text(models.kitti(96, 320, 7, False, batch=2))
    r = subprocess.run([str(exe), str(proto), "--forward"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    outs = {}
    for l in r.stdout.splitlines():
        if l.startswith("output "):
            t = l.split()
            outs[t[1]] = (float(t[t.index("sum") + 1]), float(t[t.index("abs") + 1]))
    mnet.set_precision("fp32")
    net = mnet.Net(str(proto))
    params = {}
    for i, (name, _, shapes) in enumerate(net.layers()):
        blobs = []
        for j, sh in enumerate(shapes):
            count = int(np.prod(sh))
            fan = max(count // sh[0], 1)
            scale = np.float32(3.4641) / np.sqrt(np.float32(fan)) if j == 0 else np.float32(0.2)
            blobs.append(_fill(count, i * 8 + j, scale).reshape(sh))
        if blobs:
            params[name] = blobs
    net.set_params(params)
    want = net.forward(data=_fill(2 * 3 * 96 * 320, 9999, 200.0).reshape(2, 3, 96, 320))
    assert set(outs) == set(want)
    for k, v in want.items():
        s, a = float(v.astype(np.float64).sum()), float(np.abs(v.astype(np.float64)).sum())
        assert np.isclose(outs[k][0], s, rtol=1e-6, atol=1e-6 * a) and np.isclose(outs[k][1], a, rtol=1e-6), (k, outs[k], s, a)
