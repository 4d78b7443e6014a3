"""Seeded synthetic inputs of the shapes the reference's loader produces.

Crop coordinates follow the RandomResizedCropCoord algorithm of the reference
(contrast/data/transform_coord.py:156-210: scale~U(0.08,1), log-uniform ratio in [3/4,4/3],
10-float layout [x0/(W-1), y0/(H-1), x1/(W-1), y1/(H-1), j, i, w, h, W, H], h-flip swaps
[0]<->[2], :88-92).  Flow fields are smooth random fields (coarse noise, bicubic up),
bwd = -fwd reversed in time + a small independent field, so that the forward-backward
consistency mask keeps roughly 55-70 % of pixels.
"""
import math
import random

import torch
import torch.nn.functional as F


def crop_coords(batch, width=1280, height=720, scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0),
                flip_p=0.5, seed=1234):
    rng = random.Random(seed)
    rows = []
    for _ in range(batch):
        area = height * width
        for _attempt in range(10):
            target_area = rng.uniform(*scale) * area
            aspect = math.exp(rng.uniform(math.log(ratio[0]), math.log(ratio[1])))
            w = int(round(math.sqrt(target_area * aspect)))
            h = int(round(math.sqrt(target_area / aspect)))
            if 0 < w <= width and 0 < h <= height:
                i = rng.randint(0, height - h)
                j = rng.randint(0, width - w)
                break
        else:
            w, h = width, height
            i = j = 0
        c = [float(j) / (width - 1), float(i) / (height - 1), float(j + w - 1) / (width - 1),
             float(i + h - 1) / (height - 1), float(j), float(i), float(w), float(h), float(width), float(height)]
        if rng.random() < flip_p:
            c[0], c[2] = c[2], c[0]
        rows.append(c)
    return torch.tensor(rows, dtype=torch.float32)


def flow_fields(batch, n_links, h=90, w=160, magnitude=1.5, noise=0.05, seed=1234, coarse=(9, 16)):
    """Returns (fwd, bwd), each [B,n,2,h,w] fp32 in the loader layout
    (contrast/data/dataset.py:485-495), in units of low-res pixels."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(batch * n_links, 2, coarse[0], coarse[1], generator=g)
    fwd = F.interpolate(base, size=(h, w), mode="bicubic", align_corners=True) * magnitude
    fwd = fwd.reshape(batch, n_links, 2, h, w)
    ind = torch.randn(batch * n_links, 2, coarse[0], coarse[1], generator=g)
    ind = F.interpolate(ind, size=(h, w), mode="bicubic", align_corners=True).reshape(batch, n_links, 2, h, w)
    bwd = -fwd.flip(1) + noise * magnitude * ind
    return fwd.contiguous(), bwd.contiguous()


def features(batch, channels=256, grid=7, seed=1234):
    """Projector-like features: feat~N(0,1) for both views and L2-normalised keys."""
    g = torch.Generator().manual_seed(seed)
    feat1 = torch.randn(batch, channels, grid, grid, generator=g)
    feat2 = torch.randn(batch, channels, grid, grid, generator=g)
    k1 = F.normalize(torch.randn(batch, channels, grid, grid, generator=g), dim=1)
    k2 = F.normalize(torch.randn(batch, channels, grid, grid, generator=g), dim=1)
    return feat1, feat2, k1, k2


_MOMENTUM_TWINS = (("encoder_k.", "encoder."), ("projector_instance_k.", "projector_instance."), ("projector_k.", "projector."))


def seeded_init_(module, seed=0):
    """Deterministic, construction-order-independent parameter init: every parameter is drawn from a generator
    seeded by (seed, crc32 of its state_dict name), so that two implementations with the same parameter names
    (this package's PixPro and the reference's) hold bit-identical weights whatever order they build their layers
    in.  Momentum-branch parameters (`*_k.`) use their online twin's name, i.e. start as copies (PixPro.py:280-287).
    Conv / linear weights ~ N(0, 2/fan_in), norm weights ~ 1 + 0.1 N(0,1), biases ~ 0.1 N(0,1)."""
    import zlib
    with torch.no_grad():
        for name, p in module.named_parameters():
            key = name
            for twin, online in _MOMENTUM_TWINS:
                key = key.replace(twin, online)
            g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7fffffff)
            r = torch.randn(p.shape, generator=g, dtype=torch.float32)
            if p.ndim >= 2:
                fan_in = p[0].numel()
                v = r * math.sqrt(2.0 / fan_in)
            elif name.endswith("weight"):
                v = 1.0 + 0.1 * r
            else:
                v = 0.1 * r
            p.copy_(v.to(p.device))
    return module
