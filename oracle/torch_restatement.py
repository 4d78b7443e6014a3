"""Eager-PyTorch restatement of the reference's pixel path (TEST / BASELINE INFRASTRUCTURE — not product code).

The reference (contrast/models/PixPro.py:46-247,339-363, contrast/util.py:253-357,
contrast/flow/utils/utils.py:87-89) is a sequence of torch ops; these functions restate that sequence
op for op so that it can run on the GPU box, where /root/reference does not exist.  Used by
tests/ (parity of the drop-in modules against the torch-op sequence on the same device) and by
bench.py's `torch_gpu_baseline` leg (the reference's op sequence timed on the same B200, next to the
CPU arm).  Nothing under pixpro-with-opticalflow_b200/ imports this module.
"""
import torch
import torch.nn.functional as F


def torch_featprop_fn(feat, value, gamma, clamp_value):
    """PixPro.featprop after value_transform (PixPro.py:343-363): value = value_transform(feat)."""
    N, C, H, W = feat.shape
    v = F.normalize(value, dim=1).view(N, C, -1)
    x = F.normalize(feat, dim=1).view(N, C, -1)
    att = torch.clamp(torch.bmm(x.transpose(1, 2), x), min=clamp_value)
    if gamma < 1.:
        att = att + 1e-6
    att = att ** gamma
    return torch.bmm(v, att.transpose(1, 2)).view(N, C, H, W)


def torch_featprop(model, feat):
    """PixPro.featprop with torch ops (PixPro.py:339-363)."""
    return torch_featprop_fn(feat, model.value_transform(feat), model.pixpro_p, model.pixpro_clamp_value)


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
