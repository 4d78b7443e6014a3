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


def torch_featprop(model, feat):
    """PixPro.featprop with torch ops (PixPro.py:339-363)."""
    N, C, H, W = feat.shape
    v = F.normalize(model.value_transform(feat), dim=1).view(N, C, -1)
    x = F.normalize(feat, dim=1).view(N, C, -1)
    att = torch.clamp(torch.bmm(x.transpose(1, 2), x), min=model.pixpro_clamp_value)
    if model.pixpro_p < 1.:
        att = att + 1e-6
    att = att ** model.pixpro_p
    return torch.bmm(v, att.transpose(1, 2)).view(N, C, H, W)


def torch_regression_loss(q, k, coord_q, coord_k, pos_ratio, flow=None, size=None, mask=None):
    """regression_loss with torch ops (PixPro.py:92-247; add_optical_flow :46-89 inlined)."""
    N, C, H, W = q.shape
    H_o, W_o = size
    q = q.view(N, C, -1)
    k = k.view(N, C, -1)
    xa = torch.arange(0., float(W), device=q.device).view(1, 1, -1).repeat(1, H, 1)
    ya = torch.arange(0., float(H), device=q.device).view(1, -1, 1).repeat(1, 1, W)
    qbw = ((coord_q[:, 2] - coord_q[:, 0]) / W).view(-1, 1, 1)
    qbh = ((coord_q[:, 3] - coord_q[:, 1]) / H).view(-1, 1, 1)
    kbw = ((coord_k[:, 2] - coord_k[:, 0]) / W).view(-1, 1, 1)
    kbh = ((coord_k[:, 3] - coord_k[:, 1]) / H).view(-1, 1, 1)
    qd = torch.sqrt((qbw * (W_o - 1)) ** 2 + (qbh * (H_o - 1)) ** 2)
    kd = torch.sqrt((kbw * (W_o - 1)) ** 2 + (kbh * (H_o - 1)) ** 2)
    md = torch.max(qd, kd)
    qx = ((xa + 0.5) * qbw + coord_q[:, 0].view(-1, 1, 1)) * (W_o - 1)
    qy = ((ya + 0.5) * qbh + coord_q[:, 1].view(-1, 1, 1)) * (H_o - 1)
    kx = ((xa + 0.5) * kbw + coord_k[:, 0].view(-1, 1, 1)) * (W_o - 1)
    ky = ((ya + 0.5) * kbh + coord_k[:, 1].view(-1, 1, 1)) * (H_o - 1)
    mg = None
    if flow is not None:
        gx = 2 * (qx / (W_o - 1)) - 1
        gy = 2 * (qy / (H_o - 1)) - 1
        grid = torch.stack([gx, gy], dim=-1)
        with torch.backends.cudnn.flags(enabled=False):  # ATen's native sampler, not cuDNN's (different rounding)
            fg = F.grid_sample(flow, grid, align_corners=True)
        if mask is not None:
            mg = F.grid_sample(mask.unsqueeze(1).float(), grid, mode='nearest', align_corners=True).to(torch.bool)
        qx = qx + fg[:, 0]
        qy = qy + fg[:, 1]
    dist_c = torch.sqrt((qx.view(-1, H * W, 1) - kx.view(-1, 1, H * W)) ** 2
                        + (qy.view(-1, H * W, 1) - ky.view(-1, 1, H * W)) ** 2) / md
    pos = dist_c < pos_ratio
    if mg is not None:
        pos = pos & mg.view(-1, H * W, 1)
    pf = pos.float()
    logit = torch.bmm(q.transpose(1, 2), k)
    loss = (logit * pf).sum(-1).sum(-1) / (pf.sum(-1).sum(-1) + 1e-6)
    return -2 * loss.mean(), pf.sum(-1).sum(-1)


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


def torch_flow_stage(lo_fwd, lo_bwd, alpha_1=0.01, alpha_2=0.5):
    """upflow8 + concat_flow + forward_backward_consistency with torch ops on the GPU
    (contrast/flow/utils/utils.py:87-89, contrast/util.py:253-357), for the rcp-mode comparison."""
    def up8(x):
        return 8 * F.interpolate(x, size=(8 * x.shape[2], 8 * x.shape[3]), mode='bilinear', align_corners=True)

    def ncoord(c):
        _, _, ht, wd = c.shape
        o = c.clone()
        o[:, 0] = 2 * o[:, 0] / (wd - 1) - 1
        o[:, 1] = 2 * o[:, 1] / (ht - 1) - 1
        return o

    def nflow(f):
        _, _, ht, wd = f.shape
        o = f.clone()
        o[:, 0] = 2 * o[:, 0] / (wd - 1)
        o[:, 1] = 2 * o[:, 1] / (ht - 1)
        return o

    def grid0(nb, ht, wd, dev):
        ys, xs = torch.meshgrid(torch.arange(ht, device=dev), torch.arange(wd, device=dev), indexing='ij')
        return torch.stack([xs, ys], dim=0).float().repeat(nb, 1, 1, 1)

    def concat(flows):
        num, nb, _, ht, wd = flows.shape
        if num == 1:
            return flows[0].clone()
        c0 = grid0(nb, ht, wd, flows.device)
        c1 = c0.clone()
        for f in flows:
            c1 = c1 + F.grid_sample(f, ncoord(c1).permute(0, 2, 3, 1), align_corners=True)
        return c1 - c0

    def fb(fwd, bwd):
        nb, _, ht, wd = fwd.shape
        fn, bn = nflow(fwd), nflow(bwd)
        c1 = ncoord(grid0(nb, ht, wd, fwd.device)) + fn
        m = (torch.abs(c1[:, 0]) < 1) & (torch.abs(c1[:, 1]) < 1)
        bi = F.grid_sample(bn, c1.permute(0, 2, 3, 1), align_corners=True)
        cyc = fn + bi
        a2 = alpha_2 / (torch.sqrt(torch.tensor(ht) ** 2 + torch.tensor(wd) ** 2).item())
        eps = alpha_1 * ((fn ** 2).sum(1) + (bi ** 2).sum(1)) + a2
        return m & (((cyc ** 2).sum(1) - eps) <= 0)

    B, n, _, h, w = lo_fwd.shape
    uf = up8(lo_fwd.permute(1, 0, 2, 3, 4).reshape(-1, 2, h, w)).reshape(n, B, 2, 8 * h, 8 * w)
    ub = up8(lo_bwd.permute(1, 0, 2, 3, 4).reshape(-1, 2, h, w)).reshape(n, B, 2, 8 * h, 8 * w)
    ff, fbw = concat(uf), concat(ub)
    return ff, fbw, fb(ff, fbw), fb(fbw, ff)


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
