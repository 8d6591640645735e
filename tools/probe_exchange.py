"""Where does the exchange cost come from?  One GPU, world = 1: the bench step with (a) plain detect, (b) detect_push into the
rank's own buffer (no remote store at all), (c) detect_gather over a 1-rank NCCL communicator; 30 steps each, CUDA events,
interleaved twice.  python tools/probe_exchange.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch


def main():
    from mscnn_b200 import models, net as mnet, parallel, synth
    torch.cuda.set_device(0)
    mnet.set_device(0)
    mnet.set_stream(torch.cuda.current_stream().cuda_stream)
    mnet.set_precision("fp32")
    B, H, W = 8, 768, 2560
    net = mnet.Net(models.kitti(H, W, 8, False, batch=B))
    net.set_params(synth.make_weights(net.layers()))
    img = torch.from_numpy(synth.make_images(B, H, W)).cuda()
    cfg = mnet.kitti_detect_cfg(H, W)
    cap = cfg.max_rois_per_image
    dets = torch.zeros((B, cap, 5), device="cuda")
    cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
    x = parallel.PeerExchange(B, cap, rank=0, world=1)
    comm = parallel.Comm(rank=0, world=1)
    per = parallel.payload_floats(B, cap)
    payload = torch.zeros(per, device="cuda")
    modes = {"detect": lambda: net.detect(cfg, dets.data_ptr(), cnt.data_ptr()),
             "push": lambda: net.detect_push(cfg, x),
             "nccl": lambda: net.detect_gather(cfg, comm, payload.data_ptr()),
             "none": lambda: None}

    def run(fn, steps=30):
        for _ in range(3):
            net.set_input("data", img); net.forward_only(); fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record()
        for i in range(steps):
            net.set_input("data", img); net.forward_only(); fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        t = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(steps))
        return ev[0].elapsed_time(ev[steps]) / steps, t[0], t[len(t) // 2], t[-1]

    for rep in range(2):
        for name, fn in modes.items():
            mean, lo, med, hi = run(fn)
            print(f"rep {rep} {name:7s} mean {mean:7.3f} ms  min {lo:7.3f}  median {med:7.3f}  max {hi:7.3f}", flush=True)


if __name__ == "__main__":
    main()
