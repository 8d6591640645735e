"""Micro-timings of individual layers at the bench shapes (batch 8, mscnn-8s-768) with CUDA events.
Variants are toggled through the environment (read per call by the native code)."""
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
        t = timeit(lambda: ops.conv3x3_c3_forward(x, wt, b, True, split))
        wp = ops.pack_conv_weights(wt.view(64, 27, 1, 1).contiguous(), b, split)

        def gemm_path():
            p = ops.im2col3x3_c3(x, split)
            return ops.conv_forward(p, wp, 0, relu=True)
        t2 = timeit(gemm_path)
        print(f"conv1_1 split={split}: direct {t:.3f} ms, patch-GEMM {t2:.3f} ms", flush=True)
    del x
    for name, cin, h, w, cout in [("conv1_2", 64, 768, 2560, 64), ("conv2_1", 64, 384, 1280, 128), ("conv2_2", 128, 384, 1280, 128)]:
        xin = torch.randn((B, cin, h, w), device=dev)
        wt = torch.randn((cout, cin, 3, 3), device=dev) * (2.0 / (cin * 9)) ** 0.5
        for split in (True, False):
            xp = ops.nchw_to_planes(xin, split)
            wp = ops.pack_conv_weights(wt, None, split)
            msg = f"{name} split={split}:"
            for mt in ("1", "2", "4"):
                for fat in ((False, True) if split else (False,)):
                    os.environ["MSCNN_MT"] = mt
                    if fat:
                        os.environ.pop("MSCNN_NO_FAT", None)
                    else:
                        os.environ["MSCNN_NO_FAT"] = "1"
                    t = timeit(lambda: ops.conv_forward(xp, wp, 1, relu=True))
                    msg += f"  mt{mt}{'+fat' if fat else ''} {t:.3f}"
            os.environ.pop("MSCNN_NO_FAT", None)
            os.environ.pop("MSCNN_MT", None)
            t = timeit(lambda: ops.conv_forward(xp, wp, 1, relu=True))
            msg += f"  default {t:.3f} ms"
            flops = 2.0 * B * cin * cout * 9 * h * w
            print(msg + f"  ({flops / t / 1e9:.0f} TFLOP/s algorithmic)", flush=True)
            del xp, wp
        del xin


if __name__ == "__main__":
    main()
