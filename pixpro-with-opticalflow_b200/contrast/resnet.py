"""contrast.resnet — plain PyTorch ResNet backbone (cuDNN).  OUT OF SCOPE of the hot path
(BASELINE.json north_star keeps the backbone on cuDNN); provided so that PixPro constructs
with the same state_dict keys as the reference (contrast/resnet.py): conv1, bn1,
layer{1..4}.{i}.conv{1,2,3} / bn{1,2,3} / downsample.{0,1}.  Stride sits on the 3x3 conv,
He-normal conv init, unit BN, zero-gamma on each block's last BN (contrast/resnet.py:155-173).
Only the feature-map heads ('early_return' — what PixPro uses — and 'multi_layer') are provided; every other
head and every non-default architecture keyword of the reference's ResNet raises."""
import math

import torch.nn as nn


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return self.relu(out + idt)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        return self.relu(out + idt)


class ResNet(nn.Module):
    # keyword arguments of the reference's ResNet (contrast/resnet.py) that select architectures outside the published
    # runs; accepted only at their default values — anything else raises instead of silently building a plain ResNet
    _UNSUPPORTED_DEFAULTS = dict(width=1, groups=1, width_per_group=64, deep_stem=False, avg_down=False, layer4_dilation=1,
                                 mid_dim=1024, low_dim=128)  # mid_dim / low_dim only size the pooled heads: ignored

    def __init__(self, block, layers, in_channel=3, head_type='early_return', **kw):
        super().__init__()
        for k, v in kw.items():
            if k not in self._UNSUPPORTED_DEFAULTS:
                raise TypeError(f"ResNet: unexpected keyword argument {k!r}")
            if v != self._UNSUPPORTED_DEFAULTS[k] and k not in ('low_dim', 'mid_dim'):
                raise NotImplementedError(f"ResNet: {k}={v!r} is outside the pixel-pretext scope (only the plain ResNet of the "
                                          "published runs is provided; the backbone stays on PyTorch/cuDNN)")
        if head_type not in ('early_return', 'multi_layer'):
            raise NotImplementedError(f"head_type {head_type!r}: only the feature-map heads 'early_return' (PixPro) and "
                                      "'multi_layer' are provided; the reference's pooled heads ('pass', 'mlp_head', "
                                      "'linear_head') belong to its linear-eval tooling, which is out of scope")
        self.head_type = head_type
        self.inplanes = 64
        self.conv1 = nn.Conv2d(in_channel, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = self._stage(block, 64, layers[0], 1)
        self.layer2 = self._stage(block, 128, layers[1], 2)
        self.layer3 = self._stage(block, 256, layers[2], 2)
        self.layer4 = self._stage(block, 512, layers[3], 2)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / fan))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
        last_bn = "bn3.weight" if block is Bottleneck else "bn2.weight"
        for name, p in self.named_parameters():
            if name.endswith(last_bn):
                p.data.zero_()

    def _stage(self, block, planes, n, stride):
        down = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            down = nn.Sequential(nn.Conv2d(self.inplanes, planes * block.expansion, 1, stride, bias=False),
                                 nn.BatchNorm2d(planes * block.expansion))
        blocks = [block(self.inplanes, planes, stride, down)]
        self.inplanes = planes * block.expansion
        blocks += [block(self.inplanes, planes) for _ in range(1, n)]
        return nn.Sequential(*blocks)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        c2 = self.layer1(x)
        c3 = self.layer2(c2)
        c4 = self.layer3(c3)
        c5 = self.layer4(c4)
        if self.head_type == 'multi_layer':
            return c2, c3, c4, c5
        return c5


def resnet18(**kw):
    return ResNet(BasicBlock, [2, 2, 2, 2], **kw)


def resnet34(**kw):
    return ResNet(BasicBlock, [3, 4, 6, 3], **kw)


def resnet50(**kw):
    return ResNet(Bottleneck, [3, 4, 6, 3], **kw)


def resnet101(**kw):
    return ResNet(Bottleneck, [3, 4, 23, 3], **kw)


def resnet152(**kw):
    return ResNet(Bottleneck, [3, 8, 36, 3], **kw)


# names contrast/option.py enumerates for --arch (resnet.py:5-7 of the reference); the deep-stem / wide / ResNeXt
# variants of the reference are not part of the published runs and are not provided here
__all__ = ['ResNet', 'resnet18', 'resnet34', 'resnet50', 'resnet101', 'resnet152']
