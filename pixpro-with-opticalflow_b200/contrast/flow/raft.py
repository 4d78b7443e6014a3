"""contrast.flow.raft — drop-in `RAFT` estimator (contrast/flow/raft.py:26-162 with the encoders of extractor.py and the
update operators of update.py), the flow model of the reference's non-file path (`util.calc_optical_flow`, util.py:76-103).

Division of labour, as BASELINE.json's north_star draws it: the convolutional encoders and the GRU are cuDNN work and stay
plain PyTorch modules here (built from tables below; parameter names, shapes and therefore `state_dict` keys are the
reference's, so its checkpoints load unchanged, `DataParallel`-prefixed or not); the all-pairs correlation volume, its
pyramid and the per-iteration windowed lookup — the part the reference left to an extension it does not ship
(`alt_cuda_corr`) — run on this package's kernels through `contrast.flow.corr.CorrBlock` (csrc/pp_corr.cu), and the ×8
up-sampling of the small model's prediction on `pp_upflow8`.
"""
import contextlib

import torch
import torch.nn as nn
import torch.nn.functional as F

from pixpro_b200 import ops as _ops

from .corr import CorrBlock


def _make_norm(kind, channels, groups):
    if kind == 'group':
        return nn.GroupNorm(num_groups=groups, num_channels=channels)
    if kind == 'batch':
        return nn.BatchNorm2d(channels)
    if kind == 'instance':
        return nn.InstanceNorm2d(channels)
    if kind == 'none':
        return nn.Sequential()
    raise ValueError("unknown norm_fn %r" % (kind,))


class _Block(nn.Module):
    """extractor.py:6-55 (ResidualBlock: two 3x3 convs) and :58-118 (BottleneckBlock: 1x1 / 3x3 / 1x1 at a quarter of
    the width), one class: `widths` lists the convs as (out_channels, kernel); the strided conv is the first 3x3."""

    def __init__(self, in_planes, planes, norm_fn, stride, bottleneck):
        super().__init__()
        mid = planes // 4 if bottleneck else planes
        spec = [(mid, 1), (mid, 3), (planes, 1)] if bottleneck else [(planes, 3), (planes, 3)]
        self.n = len(spec)
        self.relu = nn.ReLU(inplace=True)
        cin, strided = in_planes, False
        for i, (cout, k) in enumerate(spec, start=1):
            s = 1
            if k == 3 and not strided:
                s, strided = stride, True
            setattr(self, 'conv%d' % i, nn.Conv2d(cin, cout, kernel_size=k, padding=k // 2, stride=s))
            cin = cout
        for i, (cout, _) in enumerate(spec, start=1):
            setattr(self, 'norm%d' % i, _make_norm(norm_fn, cout, planes // 8))
        self.downsample = None
        if stride != 1:
            # the shortcut's norm is registered under its own name AND inside `downsample`, like the reference (both key
            # sets appear in a checkpoint when the norm has parameters)
            short_norm = _make_norm(norm_fn, planes, planes // 8)
            setattr(self, 'norm%d' % (self.n + 1), short_norm)
            self.downsample = nn.Sequential(nn.Conv2d(in_planes, planes, kernel_size=1, stride=stride), short_norm)

    def forward(self, x):
        y = x
        for i in range(1, self.n + 1):
            y = self.relu(getattr(self, 'norm%d' % i)(getattr(self, 'conv%d' % i)(y)))
        if self.downsample is not None:
            x = self.downsample(x)
        return self.relu(x + y)


class _Encoder(nn.Module):
    """extractor.py:121-195 (BasicEncoder) / :198-271 (SmallEncoder): 7x7 stride-2 stem, three stages of two blocks
    (strides 1, 2, 2), 1x1 output conv: 1/8 resolution."""

    def __init__(self, stem, stage_widths, bottleneck, output_dim, norm_fn, dropout):
        super().__init__()
        self.norm_fn = norm_fn
        self.norm1 = _make_norm(norm_fn, stem, 8)
        self.conv1 = nn.Conv2d(3, stem, kernel_size=7, stride=2, padding=3)
        self.relu1 = nn.ReLU(inplace=True)
        cin = stem
        for i, (w, s) in enumerate(zip(stage_widths, (1, 2, 2)), start=1):
            setattr(self, 'layer%d' % i, nn.Sequential(_Block(cin, w, norm_fn, s, bottleneck), _Block(w, w, norm_fn, 1, bottleneck)))
            cin = w
        self.conv2 = nn.Conv2d(cin, output_dim, kernel_size=1)
        self.dropout = nn.Dropout2d(p=dropout) if dropout > 0 else None
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d, nn.GroupNorm)):
                if m.weight is not None:
                    nn.init.constant_(m.weight, 1)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def forward(self, x):
        pair = isinstance(x, (tuple, list))  # both frames through the encoder as one batch
        if pair:
            nb = x[0].shape[0]
            x = torch.cat(x, dim=0)
        x = self.relu1(self.norm1(self.conv1(x)))
        x = self.conv2(self.layer3(self.layer2(self.layer1(x))))
        if self.training and self.dropout is not None:
            x = self.dropout(x)
        return torch.split(x, [nb, nb], dim=0) if pair else x


def BasicEncoder(output_dim=128, norm_fn='batch', dropout=0.0):
    return _Encoder(64, (64, 96, 128), False, output_dim, norm_fn, dropout)


def SmallEncoder(output_dim=128, norm_fn='batch', dropout=0.0):
    return _Encoder(32, (32, 64, 96), True, output_dim, norm_fn, dropout)


class FlowHead(nn.Module):
    """update.py:6-14."""

    def __init__(self, input_dim=128, hidden_dim=256):
        super().__init__()
        self.conv1 = nn.Conv2d(input_dim, hidden_dim, 3, padding=1)
        self.conv2 = nn.Conv2d(hidden_dim, 2, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.conv2(self.relu(self.conv1(x)))


class _GRU(nn.Module):
    """update.py:17-32 (ConvGRU: one 3x3 pass, gates convz / convr / convq) and :35-73 (SepConvGRU: a 1x5 pass with gates
    conv*1, then a 5x1 pass with gates conv*2)."""

    def __init__(self, hidden_dim, input_dim, separable):
        super().__init__()
        cin = hidden_dim + input_dim
        self.passes = ('1', '2') if separable else ('',)
        kernels = {'': ((3, 3), (1, 1)), '1': ((1, 5), (0, 2)), '2': ((5, 1), (2, 0))}
        for tag in self.passes:
            k, p = kernels[tag]
            for gate in 'zrq':
                setattr(self, 'conv%s%s' % (gate, tag), nn.Conv2d(cin, hidden_dim, k, padding=p))

    def forward(self, h, x):
        for tag in self.passes:
            hx = torch.cat([h, x], dim=1)
            z = torch.sigmoid(getattr(self, 'convz' + tag)(hx))
            r = torch.sigmoid(getattr(self, 'convr' + tag)(hx))
            q = torch.tanh(getattr(self, 'convq' + tag)(torch.cat([r * h, x], dim=1)))
            h = (1 - z) * h + z * q
        return h


class _MotionEncoder(nn.Module):
    """update.py:76-91 (small) / :94-112 (basic): correlation features and the current flow -> motion features."""

    def __init__(self, cor_planes, small):
        super().__init__()
        if small:
            self.convc1 = nn.Conv2d(cor_planes, 96, 1, padding=0)
            self.convf1 = nn.Conv2d(2, 64, 7, padding=3)
            self.convf2 = nn.Conv2d(64, 32, 3, padding=1)
            self.conv = nn.Conv2d(128, 80, 3, padding=1)
        else:
            self.convc1 = nn.Conv2d(cor_planes, 256, 1, padding=0)
            self.convc2 = nn.Conv2d(256, 192, 3, padding=1)
            self.convf1 = nn.Conv2d(2, 128, 7, padding=3)
            self.convf2 = nn.Conv2d(128, 64, 3, padding=1)
            self.conv = nn.Conv2d(64 + 192, 128 - 2, 3, padding=1)
        self.small = small

    def forward(self, flow, corr):
        cor = F.relu(self.convc1(corr))
        if not self.small:
            cor = F.relu(self.convc2(cor))
        flo = F.relu(self.convf2(F.relu(self.convf1(flow))))
        out = F.relu(self.conv(torch.cat([cor, flo], dim=1)))
        return torch.cat([out, flow], dim=1)


class _UpdateBlock(nn.Module):
    """update.py:115-128 (SmallUpdateBlock: no up-sampling mask) / :131-152 (BasicUpdateBlock)."""

    def __init__(self, cor_planes, hidden_dim, small):
        super().__init__()
        self.encoder = _MotionEncoder(cor_planes, small)
        if small:
            self.gru = _GRU(hidden_dim, 82 + 64, separable=False)
            self.flow_head = FlowHead(hidden_dim, hidden_dim=128)
            self.mask = None
        else:
            self.gru = _GRU(hidden_dim, 128 + hidden_dim, separable=True)
            self.flow_head = FlowHead(hidden_dim, hidden_dim=256)
            self.mask = nn.Sequential(nn.Conv2d(128, 256, 3, padding=1), nn.ReLU(inplace=True), nn.Conv2d(256, 64 * 9, 1, padding=0))

    def forward(self, net, inp, corr, flow):
        net = self.gru(net, torch.cat([inp, self.encoder(flow, corr)], dim=1))
        delta = self.flow_head(net)
        return net, (None if self.mask is None else .25 * self.mask(net)), delta  # .25: balances gradients (update.py:151)


def SmallUpdateBlock(args, hidden_dim=96):
    return _UpdateBlock(args.corr_levels * (2 * args.corr_radius + 1) ** 2, hidden_dim, small=True)


def BasicUpdateBlock(args, hidden_dim=128, input_dim=128):
    return _UpdateBlock(args.corr_levels * (2 * args.corr_radius + 1) ** 2, hidden_dim, small=False)


def coords_grid(batch, ht, wd, device=None):
    """flow/utils/utils.py:81-84: [batch, 2, ht, wd] pixel coordinates, x then y."""
    ys, xs = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing='ij')
    return torch.stack([xs, ys], dim=0).float()[None].repeat(batch, 1, 1, 1)


class RAFT(nn.Module):
    """raft.py:26-162.  `args`: a namespace with `small` (default False), `dropout` (0), `mixed_precision` (False),
    `alternate_corr` (must be False: the reference's CUDA extension for it is not shipped; this class always uses the
    kernels of csrc/pp_corr.cu)."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        for name, default in (('small', False), ('dropout', 0), ('alternate_corr', False), ('mixed_precision', False)):
            if not hasattr(args, name):
                setattr(args, name, default)
        if args.alternate_corr:
            raise NotImplementedError("alternate_corr needs the reference's unshipped alt_cuda_corr extension; the default "
                                      "CorrBlock here already runs on CUDA kernels")
        small = bool(args.small)
        self.hidden_dim, self.context_dim = (96, 64) if small else (128, 128)
        args.corr_levels, args.corr_radius = 4, (3 if small else 4)
        enc = SmallEncoder if small else BasicEncoder
        self.fnet = enc(output_dim=128 if small else 256, norm_fn='instance', dropout=args.dropout)
        self.cnet = enc(output_dim=self.hidden_dim + self.context_dim, norm_fn='none' if small else 'batch', dropout=args.dropout)
        self.update_block = (SmallUpdateBlock if small else BasicUpdateBlock)(args, hidden_dim=self.hidden_dim)

    def freeze_bn(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()

    def initialize_flow(self, img):
        """raft.py:76-84: flow = coords1 - coords0 on the 1/8 grid."""
        N, _, H, W = img.shape
        c = coords_grid(N, H // 8, W // 8, device=img.device)
        return c, c.clone()

    @staticmethod
    def upsample_flow(flow, mask):
        """raft.py:86-97: convex combination of the 3x3 neighbourhood, [N,2,H,W] -> [N,2,8H,8W]."""
        N, _, H, W = flow.shape
        mask = torch.softmax(mask.view(N, 1, 9, 8, 8, H, W), dim=2)
        nb = F.unfold(8 * flow, [3, 3], padding=1).view(N, 2, 9, 1, 1, H, W)
        up = torch.sum(mask * nb, dim=2).permute(0, 1, 4, 2, 5, 3)
        return up.reshape(N, 2, 8 * H, 8 * W)

    def _amp(self):
        if self.args.mixed_precision:
            return torch.autocast(device_type='cuda', enabled=True)
        return contextlib.nullcontext()

    def forward(self, image1, image2, iters=12, flow_init=None, upsample=True, test_mode=False):
        image1 = (2 * (image1 / 255.0) - 1.0).contiguous()
        image2 = (2 * (image2 / 255.0) - 1.0).contiguous()
        hdim, cdim = self.hidden_dim, self.context_dim
        with self._amp():
            fmap1, fmap2 = self.fnet([image1, image2])
        corr_fn = CorrBlock(fmap1.float(), fmap2.float(), num_levels=self.args.corr_levels, radius=self.args.corr_radius)
        with self._amp():
            net, inp = torch.split(self.cnet(image1), [hdim, cdim], dim=1)
            net, inp = torch.tanh(net), torch.relu(inp)
        coords0, coords1 = self.initialize_flow(image1)
        if flow_init is not None:
            coords1 = coords1 + flow_init
        predictions = []
        flow_up = None
        for _ in range(iters):
            coords1 = coords1.detach()
            corr = corr_fn(coords1)  # pp_corr_lookup: every level's window, one launch
            with self._amp():
                net, up_mask, delta = self.update_block(net, inp, corr, coords1 - coords0)
            coords1 = coords1 + delta.float()
            low = coords1 - coords0
            if up_mask is not None:
                flow_up = self.upsample_flow(low, up_mask.float())
            elif low.requires_grad:  # training the estimator: autograd through the library op (flow/utils/utils.py:87-89)
                flow_up = 8 * F.interpolate(low, size=(8 * low.shape[2], 8 * low.shape[3]), mode='bilinear', align_corners=True)
            else:
                flow_up = _ops.upflow8(low)
            predictions.append(flow_up)
        if test_mode:
            return coords1 - coords0, flow_up
        return predictions
