"""GPU: the drop-in PixPro module (contrast.models.PixPro) against a plain-PyTorch fp32
restatement of the reference's forward (featprop + regression_loss written with torch ops, as
contrast/models/PixPro.py:92-247,339-363 do), sharing the same parameters and inputs.  Loss and
all 167 parameter gradients within the fp32 tolerance; positive counts exact."""
import math
import os
import types

import pytest
import torch
import torch.distributed as dist
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle.torch_restatement import torch_featprop, torch_flow_stage, torch_regression_loss  # noqa: E402,F401
DEV = "cuda"


def pixpro_args(**kw):
    a = types.SimpleNamespace(pixpro_p=2.0, pixpro_momentum=0.99, pixpro_pos_ratio=0.7, pixpro_clamp_value=0.0,
                              pixpro_transform_layer=1, pixpro_ins_loss_weight=0.0, output_dir="/tmp",
                              num_instances=1000, batch_size=4, epochs=10, start_epoch=1, feature_dim=256,
                              head_type="early_return")
    a.__dict__.update(kw)
    return a


@pytest.fixture(scope="module")
def group():
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29551")
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo", rank=0, world_size=1)
    yield
    if dist.is_initialized():
        dist.destroy_process_group()


def torch_forward(model, im1, im2, c1, c2, f1=None, f2=None, size=(720, 1280)):
    """PixPro.forward (PixPro.py:368-439) with the torch restatements above; no EMA update."""
    pred1 = F.normalize(torch_featprop(model, model.projector(model.encoder(im1))), dim=1)
    pred2 = F.normalize(torch_featprop(model, model.projector(model.encoder(im2))), dim=1)
    with torch.no_grad():
        k1 = F.normalize(model.projector_k(model.encoder_k(im1)), dim=1)
        k2 = F.normalize(model.projector_k(model.encoder_k(im2)), dim=1)
    kw1 = dict(flow=f1[0], mask=f1[2]) if f1 is not None else {}
    kw2 = dict(flow=f2[0], mask=f2[2]) if f2 is not None else {}
    l1, n1 = torch_regression_loss(pred1, k2, c1, c2, model.pixpro_pos_ratio, size=size, **kw1)
    l2, n2 = torch_regression_loss(pred2, k1, c2, c1, model.pixpro_pos_ratio, size=size, **kw2)
    return l1 + l2, n1, n2


@pytest.mark.parametrize("use_flow", [False, True])
def test_pixpro_forward_backward_matches_torch_restatement(group, use_flow):
    from contrast import resnet, util
    from contrast.models import PixPro
    from pixpro_b200 import ops, synth
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    # the restatement runs through torch's CUDA kernels, whose `tensor / python_scalar` is x*fl(1/s):
    # compare in the matching division mode (the default 'ieee' mode is pinned to the CPU reference)
    ops.set_div_mode("rcp")
    try:
        torch.manual_seed(0)
        B = 4
        model = PixPro(resnet.resnet50, pixpro_args()).to(DEV)
        # the last BN of every block starts at gamma=0 (bag of tricks): perturb so gradients flow everywhere
        with torch.no_grad():
            for n_, p in model.named_parameters():
                if n_.endswith("bn3.weight"):
                    p.fill_(0.5)
        model.train()
        im1 = torch.randn(B, 3, 224, 224, device=DEV)
        im2 = torch.randn(B, 3, 224, 224, device=DEV)
        c1, c2 = synth.crop_coords(B, seed=1).to(DEV), synth.crop_coords(B, seed=2).to(DEV)
        coord1, coord2, f1, f2 = c1, c2, None, None
        if use_flow:
            lf, lb = synth.flow_fields(B, 2, seed=3)
            args = types.SimpleNamespace(alpha1=0.01, alpha2=0.5, use_flow_frames=False, use_flow_file=True, flow_up=True,
                                         flow_cat_norm=False, debug=False)
            data = [None] * 7
            data[5] = [torch.zeros(B), lf.to(DEV), lb.to(DEV)]
            data[6] = [torch.tensor([[720, 1280]] * B), torch.tensor([[3]] * B)]
            f1, f2 = util.apply_optical_flow(data, None, args)
            coord1, coord2 = [c1, f1], [c2, f2]
        # BatchNorm in train mode updates running stats: snapshot so both passes see the same state
        state = {k: v.clone() for k, v in model.state_dict().items()}
        loss, pos = model(im1, im2, coord1, coord2, is_update_momentum=False)
        loss.backward()
        grads = {n_: p.grad.clone() for n_, p in model.named_parameters() if p.grad is not None}
        model.zero_grad()
        model.load_state_dict(state)
        ref_loss, n1, n2 = torch_forward(model, im1, im2, c1, c2, f1, f2)
        ref_loss.backward()
        assert len(grads) == 167
        assert torch.equal(pos[0][0], n1) and torch.equal(pos[1][0], n2)
        assert abs(loss.item() - ref_loss.item()) <= 1e-5 * max(abs(ref_loss.item()), 1e-3), (loss.item(), ref_loss.item())
        gmax = max(p.grad.abs().max().item() for p in model.parameters() if p.grad is not None)
        worst, worst_name = 0.0, None
        for n_, p in model.named_parameters():
            if p.grad is None:
                continue
            denom = p.grad.abs().max().item()
            if denom < 1e-5 * gmax:
                # e.g. a bias in front of a BatchNorm: its true gradient is zero, both passes hold noise
                assert grads[n_].abs().max().item() < 1e-4 * gmax, n_
                continue
            err = (grads[n_] - p.grad).abs().max().item() / denom
            if err > worst:
                worst, worst_name = err, n_
        # the gradients run through 50 cuDNN layers whose reductions are not bit-reproducible between
        # the two passes; the pixel path itself is checked at 1e-5 in test_gpu_parity.py
        assert worst < 2e-3, (worst, worst_name)
    finally:
        ops.set_div_mode("ieee")
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.parametrize("n", [1, 3])
def test_rcp_mode_matches_torch_cuda_kernels_bitwise(n):
    """PP_DIV_RCP: the flow stage reproduces, bit for bit, what the reference computes when its
    torch ops run on the GPU through ATen's native kernels (CUDA true-divide by a scalar =
    multiply by the fp32 reciprocal; grid_sample / interpolate FMA patterns of SURVEY.md A.1)."""
    from pixpro_b200 import ops, synth
    lf, lb = synth.flow_fields(2, n, seed=40 + n)
    lf, lb = lf.to(DEV), lb.to(DEV)
    # torch's native CUDA grid_sampler_2d kernel; with cuDNN enabled torch would route this
    # bilinear/zeros/align_corners=True case to cudnnSpatialTfSamplerForward, whose rounding differs
    with torch.backends.cudnn.flags(enabled=False):
        want = torch_flow_stage(lf, lb)
    ops.set_div_mode("rcp")
    try:
        got = ops.flow_stage(lf, lb)
    finally:
        ops.set_div_mode("ieee")
    for name, g, w_ in zip(["flow_fwd", "flow_bwd", "mask_fwd", "mask_bwd"], got, want):
        assert torch.equal(g, w_), f"{name}: {(g != w_).sum().item()} of {g.numel()} differ"


def test_momentum_update_and_state_dict_roundtrip(group):
    from contrast import resnet
    from contrast.models import PixPro
    model = PixPro(resnet.resnet50, pixpro_args()).to(DEV)
    sd = model.state_dict()
    model2 = PixPro(resnet.resnet50, pixpro_args()).to(DEV)
    model2.load_state_dict(sd)
    with torch.no_grad():
        model.projector.linear2.bias.add_(1.0)
    k0 = model.k
    model._momentum_update_key_encoder()
    mom = 1. - (1. - 0.99) * (math.cos(math.pi * k0 / model.K) + 1) / 2.
    want = model2.projector_k.linear2.bias * mom + model.projector.linear2.bias * (1 - mom)
    assert torch.allclose(model.projector_k.linear2.bias, want, atol=1e-6)


@pytest.mark.parametrize("use_graph", [False, True])
def test_host_pixel_step_matches_device_path(use_graph):
    """The host-buffer entry (pinned in, pinned out; chunked copies on a side stream; optionally the
    whole step replayed from one CUDA graph) returns exactly what the device-resident ops return."""
    from pixpro_b200 import ops, synth
    from pixpro_b200.host_step import HostPixelStep
    B, C, G = 8, 256, 7
    lf, lb = synth.flow_fields(B, 1, seed=5)
    f1, f2, k1, k2 = synth.features(B, C, G, seed=6)
    c1, c2 = synth.crop_coords(B, seed=7), synth.crop_coords(B, seed=8)
    gen = torch.Generator().manual_seed(9)
    w = (torch.randn(C, C, 1, 1, generator=gen) / 16).to(DEV)
    bias = torch.zeros(C, device=DEV)
    host = {k: v.pin_memory() for k, v in dict(lo_f=lf, lo_b=lb, feat1=f1, feat2=f2, k1=k1, k2=k2, c1=c1, c2=c2).items()}
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        step = HostPixelStep(DEV, B, C, G, use_graph=use_graph)
        for _ in range(4):  # eager warm-up, capture, then replays
            out, (dw, db) = step(host, w, bias)
        ff, fb, mf, mb = ops.flow_stage(lf.to(DEV), lb.to(DEV))
        x = torch.cat([f1, f2]).to(DEV).requires_grad_(True)
        wg = w.clone().requires_grad_(True)
        p1, p2 = ops.ppm(x, F.conv2d(x, wg, bias), 2.0, 0.0, True).chunk(2)
        l12, pn, _ = ops.regression_loss_pair(p1, k2.to(DEV), c1.to(DEV), c2.to(DEV), p2, k1.to(DEV), c2.to(DEV), c1.to(DEV), 0.7,
                                              flow1=ff, flow2=fb, size=(720, 1280), mask1=mf, mask2=mb)
        (l12[0] + l12[1]).backward()
        assert torch.equal(out["pos_num"].to(DEV), pn)
        assert abs(out["loss"].item() - (l12[0] + l12[1]).item()) <= 1e-6 * max(1e-3, abs(out["loss"].item()))
        ref = x.grad.view(2, B, C, G, G).cpu()
        assert (out["d_feat"] - ref).abs().max().item() <= 1e-6 * ref.abs().max().item() + 1e-12
        assert (dw - wg.grad).abs().max().item() <= 1e-5 * wg.grad.abs().max().item()
    finally:
        torch.backends.cudnn.allow_tf32 = prev


def test_graphed_momentum_branch_matches_eager(group):
    """Opt-in CUDA-graph replay of the key branch (contrast.models.PixPro._momentum_branch_graphed): same losses as the
    eager branch step after step (EMA updates and BatchNorm running statistics included), outputs valid per call."""
    from contrast import resnet
    from contrast.models import PixPro
    from pixpro_b200 import synth
    B = 4
    torch.manual_seed(0)
    m1 = PixPro(resnet.resnet50, pixpro_args(batch_size=B)).to(DEV)
    m2 = PixPro(resnet.resnet50, pixpro_args(batch_size=B)).to(DEV)
    m2.load_state_dict(m1.state_dict())
    m2.graph_momentum_branch = True
    c1, c2 = synth.crop_coords(B, seed=1).to(DEV), synth.crop_coords(B, seed=2).to(DEV)
    gen = torch.Generator().manual_seed(3)
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for step in range(6):  # 3 eager warm-up calls, capture on the 4th, replays after
            im1 = torch.randn(B, 3, 224, 224, generator=gen).to(DEV)
            im2 = torch.randn(B, 3, 224, 224, generator=gen).to(DEV)
            with torch.no_grad():
                l1, p1 = m1(im1, im2, c1, c2)
                l2, p2 = m2(im1, im2, c1, c2)
            assert torch.equal(p1[0][0], p2[0][0]) and torch.equal(p1[1][0], p2[1][0]), step
            assert abs(l1.item() - l2.item()) <= 1e-5 * max(abs(l1.item()), 1e-3), (step, l1.item(), l2.item())
        assert m2._kgraph is not None and m2._kgraph["graph"] is not None
        for (n_, b1), (_, b2) in zip(m1.encoder_k.named_buffers(), m2.encoder_k.named_buffers()):
            assert torch.allclose(b1.float(), b2.float(), rtol=1e-5, atol=1e-6), n_
    finally:
        m2._kgraph = None
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


def test_flow_store_to_host_step_end_to_end(tmp_path):
    """SURVEY 8(f) rank 2 end to end on the GPU: per-video `.flw` stores (pixpro_b200.flowstore.write_flw) -> the drop-in
    load_flows (sliced mmap reads, contrast/data/dataset.py:341-369) -> PinnedFlowStager (pinned [B,n,2,h,w] batches) ->
    HostPixelStep (copies, flow stage, PPM, loss, backward, results back in pinned memory): loss, positive counts and feature
    gradients equal the device-resident ops fed with the same links taken straight from the tensors the stores were written
    from."""
    from pixpro_b200 import flowstore as fs
    from pixpro_b200 import ops, synth
    from pixpro_b200.host_step import HostPixelStep
    B, C, G, n = 6, 256, 7, 2
    vids = []
    for v in range(3):   # three "videos" of 12 frames at the published 90x160 flow size
        f, b = synth.flow_fields(1, 11, seed=40 + v)
        pf, pb = str(tmp_path / f"v{v}_fwd.flw"), str(tmp_path / f"v{v}_bwd.flw")
        fs.write_flw(pf, f[0])
        fs.write_flw(pb, b[0])
        vids.append((f[0], b[0], pf, pb))
    picks = [(0, 0), (0, 7), (1, 3), (1, 9), (2, 1), (2, 5)]   # (video, first link) per sample
    samples, want_f, want_b = [], [], []
    for v, s in picks:
        f, b, pf, pb = vids[v]
        bs, bn = fs.calc_bwd_idx(s, s + n, f.shape[0])
        samples.append(fs.load_flows((pf, s, s + n), (pb, s, s + n)))
        want_f.append(f[s:s + n])
        want_b.append(b[bs:bn])
    stager = fs.PinnedFlowStager(B, n, 90, 160, buffers=1)
    lo_f, lo_b = stager.collate(samples)
    assert lo_f.is_pinned() and torch.equal(lo_f, torch.stack(want_f)) and torch.equal(lo_b, torch.stack(want_b))
    f1, f2, k1, k2 = synth.features(B, C, G, seed=46)
    c1, c2 = synth.crop_coords(B, seed=47), synth.crop_coords(B, seed=48)
    gen = torch.Generator().manual_seed(49)
    w = (torch.randn(C, C, 1, 1, generator=gen) / 16).to(DEV)
    bias = torch.zeros(C, device=DEV)
    host = {k: v.pin_memory() for k, v in dict(feat1=f1, feat2=f2, k1=k1, k2=k2, c1=c1, c2=c2).items()}
    host["lo_f"], host["lo_b"] = lo_f, lo_b
    step = HostPixelStep(DEV, B, C, G, use_graph=True)
    for _ in range(4):
        out, _ = step(host, w, bias)
    ff, fb, mf, mb = ops.flow_stage(torch.stack(want_f).to(DEV), torch.stack(want_b).to(DEV))
    x = torch.cat([f1, f2]).to(DEV).requires_grad_(True)
    p1, p2 = ops.ppm(x, ops.conv1x1(x, w, bias), 2.0, 0.0, True).chunk(2)
    l12, pn, _ = ops.regression_loss_pair(p1, k2.to(DEV), c1.to(DEV), c2.to(DEV), p2, k1.to(DEV), c2.to(DEV), c1.to(DEV), 0.7,
                                          flow1=ff, flow2=fb, size=(720, 1280), mask1=mf, mask2=mb)
    (l12[0] + l12[1]).backward()
    assert torch.equal(out["pos_num"].to(DEV), pn)
    assert abs(out["loss"].item() - (l12[0] + l12[1]).item()) <= 1e-6 * max(1e-3, abs(out["loss"].item()))
    ref = x.grad.view(2, B, C, G, G).cpu()
    assert (out["d_feat"] - ref).abs().max().item() <= 1e-6 * ref.abs().max().item() + 1e-12
