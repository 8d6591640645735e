"""Micro-timings (CUDA events, batch 8, mscnn-8s-768 shapes) of the layers touched in r01f:
conv1_1 as the single tensor-core kernel, and the narrow-N fp32-faithful layers with and without the
wide-B MMA form (MSCNN_NO_WIDE=1 restores three MMAs per K step)."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from mscnn_b200 import ops


def timeit(fn, iters=5):
    for _ in range(2):
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
    dev = torch.device("cuda:0")
    B = 8
    x = (torch.randint(0, 256, (B, 3, 768, 2560), device=dev).float() - 110.0).contiguous()
    wt = torch.randn((64, 3, 3, 3), device=dev) * 0.01
    b = torch.zeros(64, device=dev)
    for split in (True, False):
        t = timeit(lambda: ops.conv1_tc_forward(x, wt, b, True, split))
        gb = B * 768 * 2560 * (12 + 128 * (2 if split else 1)) / 1e9
        print(f"conv1_1 tc split={split}: {t:.3f} ms  ({gb / t * 1e3:.0f} GB/s of algorithmic traffic)", flush=True)
    del x
    for name, cin, h, w, cout, k, pad in [("conv1_2", 64, 768, 2560, 64, 3, 1), ("conv2_1", 64, 384, 1280, 128, 3, 1),
                                          ("conv2_2", 128, 384, 1280, 128, 3, 1), ("conv3_2", 256, 192, 640, 256, 3, 1)]:
        xin = torch.randn((B, cin, h, w), device=dev)
        wt = torch.randn((cout, cin, k, k), device=dev) * (2.0 / (cin * k * k)) ** 0.5
        xp = ops.nchw_to_planes(xin, True)
        wp = ops.pack_conv_weights(wt, None, True)
        del xin
        msg = f"{name} split:"
        for wide in (False, True):
            if wide:
                os.environ.pop("MSCNN_NO_WIDE", None)
            else:
                os.environ["MSCNN_NO_WIDE"] = "1"
            t = timeit(lambda: ops.conv_forward(xp, wp, pad, relu=True))
            fl = 2.0 * B * h * w * cin * cout * k * k / 1e12
            msg += f"  {'wide' if wide else '3mma'} {t:.3f} ms ({fl / t * 1e3:.0f} TFLOP/s algorithmic)"
        os.environ.pop("MSCNN_NO_WIDE", None)
        print(msg, flush=True)
        del xp, wp
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
