"""Time the implicit-GEMM conv on the real mscnn-8s-768 layer shapes (batch 1) with CUDA
events; prints one line per layer.  Run on the GPU box: python tools/gpu_probe_conv.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from mscnn_b200 import ops

LAYERS = [  # name, Cin, H, W, Cout, k, pad
    ("conv1_2", 64, 768, 2560, 64, 3, 1),
    ("conv2_2", 128, 384, 1280, 128, 3, 1),
    ("conv3_2", 256, 192, 640, 256, 3, 1),
    ("conv4_2", 512, 96, 320, 512, 3, 1),
    ("conv5_2", 512, 48, 160, 512, 3, 1),
    ("conv6_1", 512, 24, 80, 512, 3, 1),
    ("LFCN_1_7x7", 512, 96, 320, 18, 7, 3),
]


def main():
    dev = torch.device("cuda:0")
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    for split in (False, True):
        for name, cin, h, w, cout, k, pad in LAYERS:
            x = torch.randn((batch, cin, h, w), device=dev)
            wt = torch.randn((cout, cin, k, k), device=dev) * (2.0 / (cin * k * k)) ** 0.5
            xp = ops.nchw_to_planes(x, split)
            wp = ops.pack_conv_weights(wt, None, split)
            del x
            f32 = cout < 64
            for _ in range(3):
                y = ops.conv_forward(xp, wp, pad, relu=True, out_f32=f32)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 10
            e0.record()
            for _ in range(iters):
                y = ops.conv_forward(xp, wp, pad, relu=True, out_f32=f32)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            flops = 2.0 * batch * cin * cout * k * k * h * w
            print(f"{'split' if split else 'bf16 '} {name:11s} N={batch} {ms:8.3f} ms  "
                  f"{flops / ms / 1e9:8.1f} TFLOP/s (algorithmic)", flush=True)
            del xp, wp, y


if __name__ == "__main__":
    main()
