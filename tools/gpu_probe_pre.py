"""Timing of the device pre-processing at the bench size: 8 KITTI frames 375x1242 uint8 ->
imresize (bicubic, antialiased, per-pass uint8 rounding) + BGR + mean + CHW -> 8x3x768x2560 fp32, CUDA events."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from mscnn_b200 import ops


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    B = 8
    rng = np.random.default_rng(0)
    for (ih, iw), (oh, ow), tag in [((375, 1242), (768, 2560), "KITTI upscale"), ((1366, 1024), (768, 576), "WIDER-like downscale")]:
        host = torch.from_numpy(rng.integers(0, 256, size=(B, ih, iw, 3), dtype=np.uint8)).pin_memory()
        dev = host.cuda()
        pre = ops.Preprocess((ih, iw), (oh, ow))
        out = torch.empty((B, 3, oh, ow), device="cuda")
        t_dev = timeit(lambda: pre(dev, out))
        t_host = timeit(lambda: pre(host, out))
        gb = B * (ih * iw * 3 + oh * ow * 12) / 1e9
        print(f"{tag}: {B} x {ih}x{iw} -> {oh}x{ow}: device-resident {t_dev:.3f} ms ({gb / t_dev * 1e3:.0f} GB/s in+out), "
              f"from pinned host {t_host:.3f} ms ({B * ih * iw * 3 / 1e6:.1f} MB uploaded)", flush=True)


if __name__ == "__main__":
    main()
