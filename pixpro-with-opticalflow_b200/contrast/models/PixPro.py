"""contrast.models.PixPro — the reference's model module (contrast/models/PixPro.py) with the
pixel-level pretext path (PPM, flow-guided correspondence, positive mask, masked cosine
regression loss) running in the sm_100a kernels of libpixpro_b200.so.

Same public surface as the reference: PixPro(base_encoder, args).forward(im_1, im_2, coord1,
coord2, is_update_momentum=True) -> (loss, [[pos_num, pos_mean], [pos_num, pos_mean]]);
PixPro.featprop; module-level regression_loss / add_optical_flow; state_dict keys
encoder.* projector.* encoder_k.* projector_k.* value_transform.* (checkpoint-compatible).
The backbone, projector and value_transform stay on cuDNN/cuBLAS through PyTorch.
"""
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.distributed import get_world_size

from pixpro_b200 import ops as _ops
from pixpro_b200 import optim as _optim

from .base import BaseModel


class Identity(nn.Module):
    def forward(self, input):
        return input


def conv1x1(in_planes, out_planes):
    """1x1 convolution with bias (PixPro.py:21-23)."""
    return nn.Conv2d(in_planes, out_planes, kernel_size=1, stride=1, padding=0, bias=True)


class MLP2d(nn.Module):
    """conv1x1 -> BN -> ReLU -> conv1x1 projector (PixPro.py:26-43)."""

    def __init__(self, in_dim, inner_dim=4096, out_dim=256):
        super().__init__()
        self.linear1 = conv1x1(in_dim, inner_dim)
        self.bn1 = nn.BatchNorm2d(inner_dim)
        self.relu1 = nn.ReLU(inplace=True)
        self.linear2 = conv1x1(inner_dim, out_dim)

    def forward(self, x):
        return self.linear2(self.relu1(self.bn1(self.linear1(x))))


def add_optical_flow(flow, x_grid, y_grid, size, mask=None, verbose=False):
    """PixPro.py:46-89: warp grid points (pixels of the original frame) by the flow sampled
    bilinearly at them; look the FB mask up at the nearest pixel."""
    if isinstance(flow, _ops.LazyFlow):
        flow = flow.dense()
    if isinstance(mask, _ops.LazyMask):
        mask = mask.dense()
    return _ops.add_optical_flow(flow, x_grid, y_grid, size, mask)


def _unpack_coords(coord_q, coord_k):
    """The nested-list argument convention of PixPro.py:107-124."""
    flow = size = mask = None
    if isinstance(coord_q, list):
        coord_q, flow_fwd = coord_q
        coord_k, flow_bwd = coord_k
        if isinstance(flow_fwd, list):
            flow, size, mask = flow_fwd
            if isinstance(mask, list):
                mask = mask[0]
        else:
            flow = flow_fwd
            size = flow.shape[-2:]
    return coord_q, coord_k, flow, size, mask


def regression_loss(q, k, coord_q, coord_k, pos_ratio=0.5):
    """PixPro.py:92-247.
        q, k: N * C * H * W
        coord_q, coord_k: N * 10 crop descriptors, or [coord, flow] / [coord, [flow, size, mask]]
    Returns (-2 * mean_b(masked mean of q.k), [pos_num [N], pos_mean [N]])."""
    if isinstance(coord_q, tuple):
        raise NotImplementedError("the --debug image dump of the reference (debug_utils) is out of scope")
    coord_q, coord_k, flow, size, mask = _unpack_coords(coord_q, coord_k)
    loss, pos_num, pos_mean = _ops.regression_loss(q, k, coord_q, coord_k, pos_ratio, flow=flow, size=size, mask=mask)
    return loss, [pos_num, pos_mean]


def regression_loss_pair(q1, k1, coord_q1, coord_k1, q2, k2, coord_q2, coord_k2, pos_ratio=0.5):
    """loss_1 + loss_2 of PixPro.forward (PixPro.py:429-432) in one launch.  Arguments follow
    regression_loss; returns (loss_1 + loss_2, [[pos_num, pos_mean], [pos_num, pos_mean]])."""
    if isinstance(coord_q1, tuple):
        raise NotImplementedError("the --debug image dump of the reference (debug_utils) is out of scope")
    cq1, ck1, flow1, size1, mask1 = _unpack_coords(coord_q1, coord_k1)
    cq2, ck2, flow2, size2, mask2 = _unpack_coords(coord_q2, coord_k2)
    size = size1 if size1 is not None else size2
    if q2 is None:  # q1 = [pred_1; pred_2] as one tensor: the sum and its gradient in one piece (ops._RegressionLossPairJoint)
        loss, _, pos_num, pos_mean = _ops.regression_loss_pair(q1, k1, cq1, ck1, None, k2, cq2, ck2, pos_ratio, flow1=flow1,
                                                               flow2=flow2, size=size, mask1=mask1, mask2=mask2)
        return loss, [[pos_num[0], pos_mean[0]], [pos_num[1], pos_mean[1]]]
    loss, pos_num, pos_mean = _ops.regression_loss_pair(q1, k1, cq1, ck1, q2, k2, cq2, ck2, pos_ratio, flow1=flow1,
                                                        flow2=flow2, size=size, mask1=mask1, mask2=mask2)
    return loss[0] + loss[1], [[pos_num[0], pos_mean[0]], [pos_num[1], pos_mean[1]]]


def Proj_Head(in_dim=2048, inner_dim=4096, out_dim=256):
    return MLP2d(in_dim, inner_dim, out_dim)


def Pred_Head(in_dim=256, inner_dim=4096, out_dim=256):
    return MLP2d(in_dim, inner_dim, out_dim)


class PixPro(BaseModel):
    def __init__(self, base_encoder, args):
        super().__init__(base_encoder, args)
        self.pixpro_p = args.pixpro_p
        self.pixpro_momentum = args.pixpro_momentum
        self.pixpro_pos_ratio = args.pixpro_pos_ratio
        self.pixpro_clamp_value = args.pixpro_clamp_value
        self.pixpro_transform_layer = args.pixpro_transform_layer
        self.pixpro_ins_loss_weight = args.pixpro_ins_loss_weight
        self.output_root = args.output_dir
        self.graph_momentum_branch = bool(getattr(args, "graph_momentum_branch", False))  # see _momentum_branch_graphed
        self._kgraph = None

        # online and momentum branches (PixPro.py:272-287)
        self.encoder = base_encoder(head_type='early_return')
        self.projector = Proj_Head()
        self.encoder_k = base_encoder(head_type='early_return')
        self.projector_k = Proj_Head()
        self._init_momentum_pair(self.encoder, self.encoder_k)
        self._init_momentum_pair(self.projector, self.projector_k)
        # opt-in (args.fast_sync_bn / PIXPRO_B200_FAST_SYNCBN=1): the same layers on pixpro_b200.syncbn.FastSyncBatchNorm
        # (three launches and one all-reduce per layer and direction instead of torch's ~10 launches and an all_gather)
        self.fast_sync_bn = bool(getattr(args, "fast_sync_bn", False)) or os.environ.get("PIXPRO_B200_FAST_SYNCBN", "0") == "1"
        for m in (self.encoder, self.encoder_k, self.projector, self.projector_k):
            self._sync_bn(m)

        # momentum schedule counters (PixPro.py:294-295)
        self.K = int(args.num_instances * 1. / get_world_size() / args.batch_size * args.epochs)
        self.k = int(args.num_instances * 1. / get_world_size() / args.batch_size * (args.start_epoch - 1))

        if self.pixpro_transform_layer == 0:
            self.value_transform = Identity()
        elif self.pixpro_transform_layer == 1:
            self.value_transform = conv1x1(in_planes=256, out_planes=256)
        elif self.pixpro_transform_layer == 2:
            self.value_transform = MLP2d(in_dim=256, inner_dim=256, out_dim=256)
        else:
            raise NotImplementedError

        if self.pixpro_ins_loss_weight > 0.:
            # instance branch (PixPro.py:306-319): not on the hot path, plain PyTorch
            self.projector_instance = Proj_Head()
            self.projector_instance_k = Proj_Head()
            self.predictor = Pred_Head()
            self._init_momentum_pair(self.projector_instance, self.projector_instance_k)
            for m in (self.projector_instance, self.projector_instance_k, self.predictor):
                self._sync_bn(m)
            self.avgpool = nn.AvgPool2d(7, stride=1)

    def _sync_bn(self, m):
        """PixPro.py:289-292, 315-317 (in place: every module passed is a container)."""
        nn.SyncBatchNorm.convert_sync_batchnorm(m)
        if self.fast_sync_bn:
            from pixpro_b200.syncbn import convert_fast_sync_batchnorm
            convert_fast_sync_batchnorm(m)

    @staticmethod
    def _init_momentum_pair(online, momentum):
        for p_q, p_k in zip(online.parameters(), momentum.parameters()):
            p_k.data.copy_(p_q.data)
            p_k.requires_grad = False

    def _next_momentum(self):
        """Cosine momentum schedule (PixPro.py:326-327): value for this step; advances the step counter."""
        m = 1. - (1. - self.pixpro_momentum) * (np.cos(np.pi * self.k / self.K) + 1) / 2.
        self.k = self.k + 1
        return m

    @torch.no_grad()
    def _momentum_update_key_encoder(self):
        """EMA of the key branch with the cosine momentum schedule (PixPro.py:322-337): every parameter pair in ONE
        launch (pp_ema_update).  CUDA only, like the rest of the path: a module kept on the host raises
        PixProB200Error (construction and state_dict handling work on the host; stepping does not)."""
        m = self._next_momentum()
        pairs = [(self.encoder, self.encoder_k), (self.projector, self.projector_k)]
        if self.pixpro_ins_loss_weight > 0.:
            pairs.append((self.projector_instance, self.projector_instance_k))
        qk = [(p_q.data if p_q.is_contiguous() else p_q.data.contiguous(), p_k.data)
              for online, momentum in pairs for p_q, p_k in zip(online.parameters(), momentum.parameters())]
        _optim.ema_update(qk, m, cache_key=id(self))  # validates device / dtype / layout and raises; no fallback

    def _momentum_branch(self, im_1, im_2):
        proj_1_ng = F.normalize(self.projector_k(self.encoder_k(im_1)), dim=1)
        proj_2_ng = F.normalize(self.projector_k(self.encoder_k(im_2)), dim=1)
        return proj_1_ng, proj_2_ng

    def _momentum_branch_graphed(self, im_1, im_2):
        """Opt-in (`model.graph_momentum_branch = True`): the key branch — two gradient-free encoder + projector
        passes, SyncBatchNorm collectives included — replayed from ONE CUDA graph.  Under DDP the multi-GPU step
        is bound by the host's issue rate of ~5 000 small operations (profiles/r01_s3_pretrain_ddp.txt); this
        removes the ~1 400 of the key branch.  Same kernels, same parameters (updated in place by the EMA
        launch), same running-statistics updates; the first three calls run eagerly (lazy initialisation), the
        fourth is captured.  Inputs are copied into the graph's fixed buffers; the outputs are the graph's
        buffers and are valid until the next call."""
        key = (tuple(im_1.shape), im_1.dtype, tuple(im_1.stride()), tuple(im_2.stride()), torch.is_autocast_enabled("cuda"),
               torch.get_autocast_dtype("cuda"))
        st = self._kgraph
        if st is None or st["key"] != key:
            st = self._kgraph = {"key": key, "calls": 0, "graph": None}
        if st["graph"] is None:
            st["calls"] += 1
            if st["calls"] <= 3:
                return self._momentum_branch(im_1, im_2)
            st["in"] = (im_1.clone(), im_2.clone())
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                st["out"] = self._momentum_branch(*st["in"])
            st["graph"] = graph
        st["in"][0].copy_(im_1)
        st["in"][1].copy_(im_2)
        st["graph"].replay()
        return st["out"]

    def _value(self, feat):
        """value_transform(feat) (PixPro.py:343).  The published setting (transform_layer=1, a 1x1
        conv) runs on the tcgen05 3xTF32 kernel — same parameters, fp32-accurate; Identity and the
        MLP2d variant (BatchNorm inside) stay on PyTorch/cuDNN."""
        vt = self.value_transform
        if isinstance(vt, nn.Conv2d) and vt.kernel_size == (1, 1) and feat.is_cuda:
            return _ops.conv1x1(feat, vt.weight, vt.bias)
        return vt(feat)

    def featprop(self, feat):
        """Pixel Propagation Module (PixPro.py:339-363): value transform, then the fused
        normalise / self-similarity / relu^p / propagation kernels."""
        return _ops.ppm(feat, self._value(feat), self.pixpro_p, self.pixpro_clamp_value, final_norm=False)

    def _featprop_normalized(self, feat):
        # featprop followed by F.normalize(dim=1) (PixPro.py:379-380) in one fused op; with the published single-conv value
        # transform the conv and the PPM are one autograd node (the two gradients of `feat` meet in the conv's epilogue)
        vt = self.value_transform
        if isinstance(vt, nn.Conv2d) and vt.kernel_size == (1, 1) and feat.is_cuda:
            return _ops.featprop(feat, vt.weight, vt.bias, self.pixpro_p, self.pixpro_clamp_value, final_norm=True)
        return _ops.ppm(feat, self._value(feat), self.pixpro_p, self.pixpro_clamp_value, final_norm=True)

    def regression_loss(self, x, y):
        return -2. * torch.einsum('nc, nc->n', [x, y]).mean()

    def forward(self, im_1, im_2, coord1, coord2, is_update_momentum=True):
        # online branch (PixPro.py:377-385).  The two views go through the PPM as ONE batch of 2B
        # samples (one launch instead of two: each sample is an independent thread block).
        feat_1 = self.encoder(im_1)
        proj_1 = self.projector(feat_1)
        feat_2 = self.encoder(im_2)
        proj_2 = self.projector(feat_2)
        pred_12 = self._featprop_normalized(torch.cat([proj_1, proj_2], dim=0))  # [pred_1; pred_2]

        ins = self.pixpro_ins_loss_weight > 0.
        if ins:
            def _ins(head_out):
                return F.normalize(self.avgpool(head_out).view(head_out.size(0), -1), dim=1)
            pred_instance_1 = _ins(self.predictor(self.projector_instance(feat_1)))
            pred_instance_2 = _ins(self.predictor(self.projector_instance(feat_2)))

        # momentum branch (PixPro.py:397-416)
        with torch.no_grad():
            if is_update_momentum:
                self._momentum_update_key_encoder()
            if self.graph_momentum_branch and not ins and im_1.is_cuda and self.training:
                proj_1_ng, proj_2_ng = self._momentum_branch_graphed(im_1, im_2)
            else:
                feat_1_ng = self.encoder_k(im_1)
                proj_1_ng = F.normalize(self.projector_k(feat_1_ng), dim=1)
                feat_2_ng = self.encoder_k(im_2)
                proj_2_ng = F.normalize(self.projector_k(feat_2_ng), dim=1)
                if ins:
                    proj_instance_1_ng = _ins(self.projector_instance_k(feat_1_ng))
                    proj_instance_2_ng = _ins(self.projector_instance_k(feat_2_ng))

        # pixel-level loss, both directions (PixPro.py:429-432), fused into one launch
        # (the predictions stay one tensor: no chunk / cat / 2-vector glue kernels between the loss and the PPM backward)
        loss, pos_num_list = regression_loss_pair(pred_12, proj_2_ng, coord1, coord2, None, proj_1_ng, coord2, coord1,
                                                  self.pixpro_pos_ratio)

        if ins:
            loss_instance = self.regression_loss(pred_instance_1, proj_instance_2_ng) + \
                self.regression_loss(pred_instance_2, proj_instance_1_ng)
            loss = loss + self.pixpro_ins_loss_weight * loss_instance
        return loss, pos_num_list
