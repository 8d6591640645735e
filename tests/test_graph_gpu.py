"""No layer waits for the device in the middle of a forward ("keep counts on device ... stay
graph-capturable"): BoxOutput's data-dependent row count R stays on the device for the layers behind it
(mscnn_conv_desc.dyn_n, the *_dyn entries), blob shapes are trimmed to R when the host asks (the reference's observable
shapes, box_output_layer.cpp:201), and a whole forward can be captured into a CUDA graph and replayed bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def side_stream(cuda):
    import torch
    from mscnn_b200 import net as mnet
    s = torch.cuda.Stream()
    prev = torch.cuda.current_stream()
    torch.cuda.set_stream(s)
    mnet.set_stream(s.cuda_stream)
    yield s
    torch.cuda.synchronize()
    torch.cuda.set_stream(prev)
    mnet.set_stream(prev.cuda_stream)


@pytest.mark.parametrize("kind", ["kitti8s", "kitti_cascade", "wider_cascade"])
def test_graph_replay_is_bit_identical(side_stream, kind):
    from mscnn_b200 import models, net as mnet, synth
    mnet.set_precision("fp32")
    if kind == "kitti8s":
        proto, (b, h, w) = models.kitti(192, 640, 8, False, batch=2), (2, 192, 640)
    elif kind == "kitti_cascade":
        proto, (b, h, w) = models.kitti_cascade(96, 320, batch=2), (2, 96, 320)
    else:
        proto, (b, h, w) = models.widerface_cascade(128, 192, batch=2), (2, 128, 192)
    imgs = [synth.make_images(b, h, w, first_index=i * b) for i in range(2)]
    eager = mnet.Net(proto)
    weights = synth.make_weights(eager.layers())
    eager.set_params(weights)
    want = []
    for im in imgs:
        out = eager.forward(data=im)
        want.append(({k: v.copy() for k, v in out.items()}, eager.num_proposals()))
    assert want[0][1] != want[1][1], "the two batches should give different proposal counts"
    net = mnet.Net(proto)
    net.set_params(weights)
    net.set_graph(True)
    replays = 0
    for step, which in enumerate([0, 0, 1, 0, 1, 1]):
        out = net.forward(data=imgs[which])
        replays += int(net.graph_replayed())
        ref, rows = want[which]
        assert net.num_proposals() == rows, (step, net.num_proposals(), rows)
        assert net.blob_shape("proposals")[0] == max(rows, 1)
        for k in ref:
            assert out[k].shape == ref[k].shape, (step, k, out[k].shape, ref[k].shape)
            assert np.array_equal(out[k], ref[k]), (step, k)
    assert replays >= 4, replays          # the first forward is eager (allocations), the others are graph launches
    # a parameter update invalidates the graph: the next forward runs eagerly again and sees the new weights
    w2 = [a.copy() for a in weights["cls_pred"]]
    w2[1][:] += 0.5
    net.set_params({"cls_pred": w2})
    eager.set_params({"cls_pred": w2})
    a, b2 = net.forward(data=imgs[0]), eager.forward(data=imgs[0])
    assert not net.graph_replayed()
    key = "cls_pred" if "cls_pred" in a else sorted(a)[0]
    assert np.array_equal(a[key], b2[key])
    if key == "cls_pred":
        assert not np.array_equal(a[key], want[0][0][key])


def test_rows_are_deferred_and_shapes_follow_the_data(cuda):
    """forward_only() returns with the forward queued; the blobs behind BoxOutput report the true R once asked, for
    batches with different R in a row, and a partial forward fed with an injected ROI list uses the injected count."""
    from mscnn_b200 import models, net as mnet, synth
    mnet.set_precision("fp32")
    b, h, w = 2, 96, 320
    net = mnet.Net(models.kitti(h, w, 7, False, batch=b))
    net.set_params(synth.make_weights(net.layers()))
    rows = []
    for i in range(3):
        net.set_input("data", synth.make_images(b, h, w, first_index=i * b))
        net.forward_only()
        r = net.num_proposals()
        rows.append(r)
        assert net.blob_shape("proposals") == (r, 5, 1, 1)
        assert net.blob_shape("cls_pred")[0] == r and net.blob_shape("fc6")[0] == r
        assert net.num_proposals(0) + net.num_proposals(1) == r
    assert len(set(rows)) > 1
    full = {k: net.blob(k) for k in ("cls_pred", "bbox_pred")}
    props = net.blob("proposals")
    keep = props[: rows[-1] // 2]                      # inject HALF of the ROIs: the head must process exactly those
    for name in net.blob_names:
        if name.startswith("proposals_proposals_0_split_"):
            net.set_input(name, keep)
    net.forward_only(start="roi_pool_org")
    for k in full:
        got = net.blob(k)
        assert got.shape[0] == len(keep)
        assert np.array_equal(got, full[k][: len(keep)]), k


def test_partial_forward_may_start_inside_a_fused_group(cuda):
    """Net::ForwardFromTo(start, end) is part of the mirrored API (net.cpp:544-555).  A range that starts at a layer the
    fusion pass folded into another one (pool1 lives in conv1_2's epilogue, roi_pool_ctx is pooled by roi_pool_org's
    launch) is widened back to that layer, and a mark left by a range that ended between the two never leaks into the
    next call: every such partial forward reproduces the full forward bit for bit."""
    from mscnn_b200 import models, net as mnet, synth
    mnet.set_precision("fp32")
    b, h, w = 2, 96, 320
    net = mnet.Net(models.kitti(h, w, 8, False, batch=b))
    net.set_params(synth.make_weights(net.layers()))
    want = {k: v.copy() for k, v in net.forward(data=synth.make_images(b, h, w)).items()}
    pool3 = net.blob("pool3")
    for start in ("pool1", "pool3", "roi_pool_ctx", "roi_pool_org"):
        net.forward_only(start=start)
        got = {k: net.blob(k) for k in want}
        for k in want:
            assert np.array_equal(got[k], want[k]), (start, k)
    net.forward_only(end="conv3_3")          # ends between conv3_3 (which pools) and pool3
    net.forward_only(start="conv4_1")        # must NOT find a stale "already pooled" mark anywhere
    assert np.array_equal(net.blob("pool3"), pool3)
    for k in want:
        assert np.array_equal(net.blob(k), want[k]), k
