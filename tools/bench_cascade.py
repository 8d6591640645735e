"""Measurement of the cascade deploy nets -- NOT the driver's bench line
(bench.py measures BASELINE.json's metric on mscnn-8s-768); same protocol, one GPU:

    python tools/bench_cascade.py [--net kitti|wider] [--batch 8] [--steps 10] [--warmup 3]

One step = one forward of the batch + the stage-3 final-detection post-process, inputs resident in HBM,
timed with CUDA events; e2e = pinned-host input + D2H of the detections inside the step; per-layer CUDA-event
times give the conv+fc share and its tensor roofline fraction; the reference's own CPU code (oracle/_ref)
runs a bounded sample beside it.  Prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402  (conv_flops, peaks, ClockSampler)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--net", default="kitti", choices=["kitti", "wider"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    import torch
    from mscnn_b200 import capi, models, net as mnet, synth

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    mnet.set_device(0)
    mnet.set_stream(torch.cuda.current_stream().cuda_stream)
    B = args.batch
    if args.net == "kitti":
        H, W, ncls, overlap, prob_blob = 576, 1920, 5, 0.5, None
        proto, sample_hw = models.kitti_cascade(H, W, batch=B), (192, 640)
        workload = "cascade-mscnn-7s-576-2x KITTI-car forward + stage-3 post-process, 3x576x1920 synthetic"
    else:
        H, W, ncls, overlap, prob_blob = 768, 1024, 2, 0.3, "cls_prob_3rd_avg"
        proto, sample_hw = models.widerface_cascade(H, W, batch=B), (192, 256)
        workload = "cascade-mscnn-12s-align WIDER-face forward + averaged stage-3 post-process, 3x768x1024 synthetic"
    net = mnet.Net(proto)
    net.set_params(synth.make_weights(net.layers()))
    host_img = torch.from_numpy(synth.make_images(B, H, W)).pin_memory()
    dev_img = host_img.to(dev)
    cfg = capi.DetectCfg()
    cfg.num_cls, cfg.cls_id, cfg.nms_overlap = ncls, 2, overlap
    cfg.ratio_h = cfg.ratio_w = 1.0
    cfg.org_h, cfg.org_w = float(H), float(W)
    cfg.max_rois_per_image = 3000
    cap = cfg.max_rois_per_image
    dets = torch.zeros((B, cap, 5), device=dev)
    cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    host_dets = torch.zeros((B, cap, 5)).pin_memory()
    host_cnt = torch.zeros(B, dtype=torch.int32).pin_memory()

    def step(src):
        net.set_input("data", src)
        net.forward_only()
        net.detect_cascade(cfg, dets.data_ptr(), cnt.data_ptr(), stage="3rd", cls_prob=prob_blob)

    def step_e2e():
        step(host_img)
        host_dets.copy_(dets, non_blocking=True)
        host_cnt.copy_(cnt, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        n0 = capi.lib().mscnn_kernel_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps, (capi.lib().mscnn_kernel_launch_count() - n0) // args.steps

    out = {"workload": workload, "batch": B, "steps": args.steps, "warmup": args.warmup, "n_gpus": 1,
           "metric": "images/sec", "unit": "images/s", "data": "synthetic (mscnn_b200/synth.py, seed 1706)"}
    pk = bench.peaks()
    sampler = bench.ClockSampler(0)
    for mode in ("fp32", "bf16"):
        mnet.set_precision(mode)
        if mode == "fp32":
            sampler.start()
        ms, launches = timed(lambda: step(dev_img))
        clocks = sampler.stop() if mode == "fp32" else None
        ms_e2e, _ = timed(step_e2e)
        lt, lt2 = net.time_layers(), net.time_layers()
        types = dict(zip(net.layer_names, net.layer_types))
        conv_ms = sum(min(lt[k], lt2[k]) for k in lt if types[k] in ("Convolution", "InnerProduct"))
        all_ms = sum(min(lt[k], lt2[k]) for k in lt)
        flops, _ = bench.conv_flops(net, B)
        r = {"ms_per_step": ms, "value": B / (ms / 1e3), "e2e_ms_per_step": ms_e2e, "e2e_value": B / (ms_e2e / 1e3),
             "kernel_launches_per_step": launches, "proposals_per_image": net.num_proposals() / B,
             "detections_per_image": float(cnt.float().mean().item()),
             "conv_fc_ms": conv_ms, "conv_fc_share": conv_ms / all_ms, "algorithmic_tflop_per_step": flops / 1e12,
             "roofline": {"bound": "tensor", "achieved": flops / (conv_ms / 1e3) / 1e12, "peak": pk["tflops"],
                          "unit": "TFLOP/s", "frac": flops / (conv_ms / 1e3) / 1e12 / pk["tflops"]},
             "top_layers_ms": sorted(((round(min(lt[k], lt2[k]), 3), k) for k in lt), reverse=True)[:8],
             "by_type_ms": {}}
        for k in lt:
            r["by_type_ms"][types[k]] = round(r["by_type_ms"].get(types[k], 0.0) + min(lt[k], lt2[k]), 3)
        if clocks:
            r["clocks"] = clocks
        out["fp32_faithful" if mode == "fp32" else "bf16"] = r
    mnet.set_precision("fp32")
    if not args.no_cpu_baseline:
        from oracle import ref
        if ref.available():
            sh, sw = sample_hw
            gen = models.kitti_cascade if args.net == "kitti" else models.widerface_cascade
            rnet = ref.RefNet(gen(sh, sw, batch=1), is_path=False)
            layers = [(n, t, rnet.param_shapes(n)) for n, t in zip(rnet.layer_names, rnet.layer_types)]
            rnet.set_params(synth.make_weights(layers))
            rnet.set_blob("data", synth.make_images(1, sh, sw))
            rnet.forward()
            t0 = time.perf_counter()
            rnet.forward()
            dt = time.perf_counter() - t0
            frac = sh * sw / float(H * W)
            out["cpu_baseline"] = {"value": frac / dt, "unit": "images/s", "cores": ref.blas_threads(), "kind": "reference",
                                   "sample": f"1 forward of a 3x{sh}x{sw} image (= {frac:.4f} of 3x{H}x{W}) through the same net "
                                             f"with the reference's CPU layers (oracle/_ref, {ref.blas_backend()}); {dt:.2f} s"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
