#!/usr/bin/env python
"""Benchmark of the MS-CNN detection forward path (BASELINE.json: images/sec, mscnn-8s-768
KITTI-car forward, batch 8 per B200, 3x768x2560 synthetic input; proposals/sec reported beside it).

    python bench.py --gpus 1 --steps K --warmup W            # this framework, one process per GPU
    torchrun --nproc-per-node N bench.py --gpus N ...        # image-parallel, weak scaling
    python bench.py --impl reference ...                     # the reference's CPU code on host cores

A "step" = one forward of one batch of 8 images per GPU through the whole path: conv trunk,
proposal heads, BoxOutput (decode + top-2000 + NMS), ROI pooling, detection head, final-detection
post-process; for N > 1 followed by one NCCL all-gather of the final boxes.

  value  : images/s, whole job, inputs already resident in HBM (fp32-faithful split-bf16 path, the
           path that meets the 1e-3 parity gate); `bf16` carries the same measurement for the plain
           bf16 tensor-core path (config 3 of BASELINE.json asks for both).
  e2e    : same metric through the public API (mscnn_b200.net.Net) with HOST buffers: pinned-host
           -> device copy of the batch and device -> host copy of the detections inside every step.
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md section "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

NET_H, NET_W, BATCH = 768, 2560, 8
WORKLOAD = "mscnn-8s-768 KITTI-car forward, batch 8 per GPU, 3x768x2560 synthetic (BASELINE.json configs[2])"
CPU_SAMPLE_H, CPU_SAMPLE_W = 192, 640          # 1/16 of the pixels of one 768x2560 image


def peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))),
                "tflops_burst": float(d.get("bf16_tflops", 1590.0)), "hbm_gbs": float(d.get("hbm_gbs", 6650.0)),
                "source": "MEASURED_PEAKS.json (sustained cuBLAS bf16: kernel timed inside a long step)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median of the upper half = clocks while the GPU is loaded
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


def conv_flops(net, n_images: int) -> tuple[float, int]:
    """Algorithmic FLOPs of one forward (2*Cin*Cout*kh*kw*Hout*Wout*N per Convolution, 2*M*N*K per
    InnerProduct) using the shapes of the LAST forward (R is data dependent)."""
    total, launches = 0.0, 0
    for name, ltype, shapes in net.layers():
        if ltype == "Convolution":
            co, ci, kh, kw = shapes[0]
            n, _, ho, wo = net.blob_shape(name)
            total += 2.0 * ci * co * kh * kw * ho * wo * n
            launches += 1
        elif ltype == "InnerProduct":
            no, k = shapes[0]
            total += 2.0 * net.blob_shape(name)[0] * no * k
            launches += 1
    return total, launches


def run_reference(args) -> None:
    """--impl reference: the reference's own CPU implementation (oracle/_ref = its layer sources
    compiled verbatim, else the oracle port is reported as unavailable), all host threads, on a
    bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref
    from mscnn_b200 import models, synth
    base = {"impl": "reference", "metric": "images/sec", "unit": "images/s", "higher_is_better": True,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "scaling": "weak", "dtype": "f32",
            "data": "synthetic", "vs_baseline": None,
            "config": {"workload": WORKLOAD, "global_batch": BATCH * args.gpus, "parallelism": f"dp{args.gpus}"}}
    if not ref.available():
        print(json.dumps({**base, "unavailable": "oracle/_ref/libmscnn_ref.so was not built (needs /root/reference at build time)"}))
        return
    frac = (CPU_SAMPLE_H * CPU_SAMPLE_W) / float(NET_H * NET_W)
    net = ref.RefNet(models.kitti(CPU_SAMPLE_H, CPU_SAMPLE_W, 8, False, batch=1), is_path=False)
    layers = [(n, t, net.param_shapes(n)) for n, t in zip(net.layer_names, net.layer_types)]
    net.set_params(synth.make_weights(layers))
    net.set_blob("data", synth.make_images(1, CPU_SAMPLE_H, CPU_SAMPLE_W))
    for _ in range(max(1, min(args.warmup, 1))):
        net.forward()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        net.forward()
    dt = (time.perf_counter() - t0) / args.steps
    value = frac / dt
    sample = (f"per step: 1 synthetic image of 3x{CPU_SAMPLE_H}x{CPU_SAMPLE_W} (= {frac:.4f} of a 3x{NET_H}x{NET_W} image) "
              f"through the full mscnn-8s net, Caffe CPU mode, value = {frac:.4f} / seconds; {ref.blas_backend()}")
    line = {**base, "value": value, "ms_per_step": dt * 1e3,
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": ref.blas_threads(), "kind": "reference",
                             "sample": sample},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def cpu_baseline(seconds_budget: float = 25.0) -> dict:
    from oracle import ref
    from mscnn_b200 import models, synth
    if not ref.available():
        return {"value": None, "unit": "images/s", "cores": 0, "kind": "reference",
                "sample": "oracle/_ref not built on this machine"}
    frac = (CPU_SAMPLE_H * CPU_SAMPLE_W) / float(NET_H * NET_W)
    net = ref.RefNet(models.kitti(CPU_SAMPLE_H, CPU_SAMPLE_W, 8, False, batch=1), is_path=False)
    layers = [(n, t, net.param_shapes(n)) for n, t in zip(net.layer_names, net.layer_types)]
    net.set_params(synth.make_weights(layers))
    net.set_blob("data", synth.make_images(1, CPU_SAMPLE_H, CPU_SAMPLE_W))
    t0 = time.perf_counter()
    net.forward()                       # warm-up forward (caffe time does one, tools/caffe.cpp:359-362)
    warm = time.perf_counter() - t0
    reps = max(1, min(3, int(seconds_budget / max(warm, 1e-3)) - 1))
    t0 = time.perf_counter()
    for _ in range(reps):
        net.forward()
    dt = (time.perf_counter() - t0) / reps
    return {"value": frac / dt, "unit": "images/s", "cores": ref.blas_threads(), "kind": "reference",
            "sample": (f"{reps} forward(s) of 1 synthetic 3x{CPU_SAMPLE_H}x{CPU_SAMPLE_W} image (= {frac:.4f} of one "
                       f"3x{NET_H}x{NET_W} image) through the full mscnn-8s net with the reference's own CPU layers "
                       f"(compiled verbatim, Caffe CPU mode, {ref.blas_backend()}, {os.cpu_count()} host cpus); "
                       f"{dt:.2f} s per sample; images/s = {frac:.4f} / s"),
            "seconds_per_sample": dt}


def parallel_first_index(rank: int, per_gpu_batch: int) -> int:
    from mscnn_b200 import parallel
    return parallel.shard_range(rank, per_gpu_batch)[0]


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mscnn_b200", choices=["mscnn_b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="images per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bf16", action="store_true", help="skip the plain-bf16 measurement")
    ap.add_argument("--profile-bf16", action="store_true",
                    help="profiling aid: run ONLY plain-bf16 forwards and print nothing (for ncu captures)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "mscnn_b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from mscnn_b200 import capi, models, net as mnet, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its banner ("NCCL version ...") to stdout when NCCL_DEBUG >= VERSION: keep stdout = the JSON line
        # (NCCL honours NCCL_DEBUG_FILE only above the VERSION level, so VERSION is raised to WARN)
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    mnet.set_device(local)
    mnet.set_stream(torch.cuda.current_stream().cuda_stream)
    B = args.batch

    net = mnet.Net(models.kitti(NET_H, NET_W, 8, False, batch=B))
    net.set_params(synth.make_weights(net.layers()))
    first = parallel_first_index(rank, B)
    host_img = torch.from_numpy(synth.make_images(B, NET_H, NET_W, first_index=first)).pin_memory()
    dev_img = host_img.to(dev, non_blocking=True)
    cfg = mnet.kitti_detect_cfg(NET_H, NET_W)
    cap = cfg.max_rois_per_image
    dets = torch.zeros((B, cap, 5), device=dev)
    cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    host_dets = torch.zeros((B, cap, 5)).pin_memory()
    host_cnt = torch.zeros(B, dtype=torch.int32).pin_memory()
    from mscnn_b200 import parallel
    gbuf = parallel.GatherBuffers(world, B, cap, dev) if world > 1 else None

    def gather():
        if world > 1:   # the path's only exchange: final boxes of every rank
            parallel.all_gather_detections(dets, cnt, gbuf)

    def step_resident():
        net.set_input("data", dev_img)          # D2D into the net's input blob (inputs resident in HBM)
        net.forward_only()
        net.detect(cfg, dets.data_ptr(), cnt.data_ptr())
        gather()

    def step_e2e_serial():
        net.set_input("data", host_img)         # pinned host -> device, async on the net stream
        net.forward_only()
        net.detect(cfg, dets.data_ptr(), cnt.data_ptr())
        gather()
        host_dets.copy_(dets, non_blocking=True)
        host_cnt.copy_(cnt, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def step_e2e():
        # Software-pipelined over steps, as a throughput deployment feeds the net: every step still does one
        # H2D upload of a full input batch from pinned host memory and one D2H read of its detections, but the
        # upload is the NEXT step's input, issued on the copy stream as soon as this step's forward has been
        # queued (it starts on the device once conv1_1 of this step has consumed the blob), so it overlaps the
        # ROI head instead of preceding the trunk.  The first timed step's input was uploaded by the last
        # warm-up step; K timed steps contain exactly K uploads and K downloads.
        net.forward_only()
        net.detect(cfg, dets.data_ptr(), cnt.data_ptr())
        gather()
        net.set_input_async("data", host_img)
        host_dets.copy_(dets, non_blocking=True)
        host_cnt.copy_(cnt, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    # the MATLAB driver's own entry: ORIGINAL uint8 frames (KITTI: 375 x 1242) -> imresize + BGR + mean + CHW on
    # the device -> forward -> detections; 11 MB of H2D per step instead of 189 MB
    from mscnn_b200 import ops as mops
    ORG_H, ORG_W = 375, 1242
    pre = mops.Preprocess((ORG_H, ORG_W), (NET_H, NET_W))
    rng_u8 = np.random.default_rng(1706 + first)
    host_u8 = torch.from_numpy(rng_u8.integers(0, 256, size=(B, ORG_H, ORG_W, 3), dtype=np.uint8)).pin_memory()

    def step_e2e_images():
        net.set_input_images("data", pre, host_u8)
        net.forward_only()
        net.detect(cfg, dets.data_ptr(), cnt.data_ptr())
        gather()
        host_dets.copy_(dets, non_blocking=True)
        host_cnt.copy_(cnt, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = capi.lib().mscnn_kernel_launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        launched = capi.lib().mscnn_kernel_launch_count() - n0
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1), float(launched)], device=dev)
        if world > 1:
            both = ms.clone()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(both, op=dist.ReduceOp.SUM)
            ms[1] = both[1]
        timed.launched = int(ms[1].item())     # kernels launched by libmscnn_b200 in the timed region, all ranks
        return float(ms[0].item())

    if args.profile_bf16:
        mnet.set_precision("bf16")
        for _ in range(args.warmup + args.steps):
            step_resident()
        torch.cuda.synchronize()
        return
    results = {}
    sampler = ClockSampler(local) if rank == 0 else None
    for mode in (["fp32"] if args.no_bf16 else ["fp32", "bf16"]):
        mnet.set_precision(mode)
        if mode == "fp32" and sampler:
            sampler.start()
        ms = timed(step_resident, args.steps, args.warmup)
        launched = timed.launched
        clocks = sampler.stop() if (mode == "fp32" and sampler) else None
        props = torch.tensor([float(net.num_proposals())], device=dev)
        if world > 1:
            dist.all_reduce(props)
        ms_e2e_serial = timed(step_e2e_serial, args.steps, args.warmup)
        net.set_input_async("data", host_img)   # prologue of the pipelined loop
        ms_e2e = timed(step_e2e, args.steps, args.warmup)
        ms_e2e_images = timed(step_e2e_images, args.steps, args.warmup) if mode == "fp32" else None
        net.set_input("data", dev_img)
        # per-layer device times for the roofline: two extra forwards with CUDA events per layer
        mnet.set_precision(mode)
        lt = net.time_layers()
        lt2 = net.time_layers()
        types = dict(zip(net.layer_names, net.layer_types))
        conv_ms = sum(min(lt[k], lt2[k]) for k in lt if types[k] in ("Convolution", "InnerProduct"))
        all_ms = sum(min(lt[k], lt2[k]) for k in lt)
        flops, conv_launches = conv_flops(net, B)
        results[mode] = dict(ms=ms, ms_e2e=ms_e2e, ms_e2e_serial=ms_e2e_serial, ms_e2e_images=ms_e2e_images, props=float(props.item()), conv_ms=conv_ms, all_ms=all_ms,
                             flops=flops, conv_launches=conv_launches, clocks=clocks, launched=launched,
                             top=sorted(((min(lt[k], lt2[k]), k) for k in lt), reverse=True)[:6],
                             layers={k: round(min(lt[k], lt2[k]), 3) for k in lt if min(lt[k], lt2[k]) >= 0.02})
    mnet.set_precision("fp32")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    r = results["fp32"]
    total_images = B * world * args.steps
    value = total_images / (r["ms"] / 1e3)
    e2e_value = total_images / (r["ms_e2e"] / 1e3)
    achieved = r["flops"] / (r["conv_ms"] / 1e3) / 1e12
    line = {
        "metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms"] / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": "bf16x3 (fp32-faithful: 3-term bf16 split on tcgen05, fp32 accumulate in TMEM)",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": B * world, "per_gpu_batch": B,
                   "parallelism": f"dp{world} image-parallel, one NCCL all-gather of final boxes",
                   "l2": "inputs larger than L2 (189 MB image batch + >10 GB of activations per step vs 126 MB L2)",
                   "weights": "seeded synthetic (mscnn_b200/synth.py), seed 1706"},
        "proposals_per_sec": r["props"] * args.steps / (r["ms"] / 1e3),
        "proposals_per_image": r["props"] / (B * world),
        "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": r["ms_e2e"] / args.steps,
                "h2d_bytes_per_step": B * 3 * NET_H * NET_W * 4 * world,
                "d2h_bytes_per_step": (B * cap * 5 * 4 + B * 4) * world,
                "pipelined": True,   # the upload inside step k is step k+1's batch; serial_* below is the strict order
                "serial_value": total_images / (r["ms_e2e_serial"] / 1e3),
                "serial_ms_per_step": r["ms_e2e_serial"] / args.steps,
                "api": "mscnn_b200.net.Net: forward_only / detect / set_input_async(pinned host, next step's batch, "
                       "copy stream) / D2H of detections + sync, software-pipelined over steps (one full-batch upload "
                       "and one download inside every timed step); serial_* = set_input / forward / detect / D2H "
                       "strictly in order on one stream"},
        # same loop fed with ORIGINAL uint8 frames: upload + device pre-processing (imresize bicubic/antialias, BGR,
        # mean, CHW) + forward + detect + download, strictly serial on the net stream (random uint8 frames, so the
        # proposal count differs from the headline workload)
        "e2e_images": {"value": total_images / (r["ms_e2e_images"] / 1e3), "unit": "images/s",
                       "ms_per_step": r["ms_e2e_images"] / args.steps,
                       "h2d_bytes_per_step": B * 375 * 1242 * 3 * world,
                       "d2h_bytes_per_step": (B * cap * 5 * 4 + B * 4) * world,
                       "api": "Net.set_input_images(uint8 375x1242 frames, ops.Preprocess) / forward_only / detect / D2H"},
        # counted by the library itself (mscnn_kernel_launch_count) inside the timed resident region, all ranks
        "gpu_launches": r["launched"],
        "roofline": {"bound": "tensor", "kernel": "conv_igemm_kernel<BLOCK_N> + conv_c3_tc_kernel (all Convolution + InnerProduct layers)",
                     "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"],
                     # mean dram__bytes_read + dram__bytes_write per launch of the convolution kernels
                     # (conv_igemm_kernel<*> + conv_c3_tc_kernel, 378 launches = 14 forwards) from the committed ncu
                     # launch list profiles/r01i_launches.csv; per-layer `ncu --set full` captures (conv3_2: 1.02 GB +
                     # 0.97 GB = its algorithmic 2.01 GB of planes in + out) in profiles/r01h_summary.md / r01i_summary.md
                     "traffic": 1.266e+09, "traffic_source": "profiles/r01i_launches.csv (mean over the step's conv launches)",
                     "algorithmic_gflop_per_step": r["flops"] / 1e9,
                     "launches_per_step": r["conv_launches"],
                     "avg_launch_ms": r["conv_ms"] / r["conv_launches"],
                     "conv_share_of_step": r["conv_ms"] / r["all_ms"],
                     "executed_tensor_flops_factor": 3,
                     "peak_source": pk["source"],
                     "how": "algorithmic FLOPs of one step / sum of per-layer CUDA-event times of the conv+fc layers "
                            "(two timed forwards after the timed region, min per layer, same stream)"},
        "clocks": r["clocks"],
        "top_layers_ms": [[k, round(v, 3)] for v, k in r["top"]],
        "layers_ms": r["layers"],
    }
    if "bf16" in results:
        b = results["bf16"]
        ach_b = b["flops"] / (b["conv_ms"] / 1e3) / 1e12
        line["bf16"] = {"value": total_images / (b["ms"] / 1e3), "unit": "images/s", "ms_per_step": b["ms"] / args.steps,
                        "e2e_value": total_images / (b["ms_e2e"] / 1e3),
                        "e2e_serial_value": total_images / (b["ms_e2e_serial"] / 1e3),
                        "roofline": {"achieved": ach_b, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach_b / pk["tflops"]},
                        "note": "plain bf16 conv path: fails the 1e-3 parity gate (bf16 rounding of activations), "
                                "reported as BASELINE.json config 3 asks; proposals/image differ accordingly",
                        "proposals_per_image": b["props"] / (B * world),
                        "top_layers_ms": [[k, round(v, 3)] for v, k in b["top"]], "layers_ms": b["layers"]}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
