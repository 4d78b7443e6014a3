"""CPU: host-side logic of the drop-in mirror (no kernels run): module construction and
checkpoint key compatibility, argument unpacking, sub-chain enumeration, EMA schedule, error
behaviour, and the multi-rank bookkeeping of bench.py under a 2-process gloo group."""
import math
import os
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def pixpro_args(**kw):
    a = types.SimpleNamespace(pixpro_p=2.0, pixpro_momentum=0.99, pixpro_pos_ratio=0.7, pixpro_clamp_value=0.0,
                              pixpro_transform_layer=1, pixpro_ins_loss_weight=0.0, output_dir="/tmp",
                              num_instances=1000, batch_size=4, epochs=10, start_epoch=1, feature_dim=256,
                              head_type="early_return")
    a.__dict__.update(kw)
    return a


@pytest.fixture(scope="module")
def gloo1():
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29541")
        dist.init_process_group("gloo", rank=0, world_size=1)
    yield
    if dist.is_initialized():
        dist.destroy_process_group()


def test_pixpro_constructs_with_reference_state_dict_keys(gloo1):
    from contrast import resnet
    from contrast.models import PixPro
    m = PixPro(resnet.resnet50, pixpro_args())
    keys = list(m.state_dict().keys())
    prefixes = {k.split(".")[0] for k in keys}
    assert prefixes == {"encoder", "projector", "encoder_k", "projector_k", "value_transform"}
    assert "value_transform.weight" in keys and "value_transform.bias" in keys
    assert "encoder.layer4.2.bn3.weight" in keys and "projector.linear2.bias" in keys
    trainable = [p for p in m.parameters() if p.requires_grad]
    # 167 trainable tensors, 33 023 552 fp32 values (encoder + projector + value_transform)
    assert len(trainable) == 167
    assert sum(p.numel() for p in trainable) == 33_023_552
    # momentum branch starts as a copy and takes no gradient
    assert all(not p.requires_grad for p in m.encoder_k.parameters())
    assert torch.equal(m.encoder.conv1.weight, m.encoder_k.conv1.weight)
    # SyncBN conversion happened in place on the children (PixPro.py:289-292)
    assert isinstance(m.encoder.bn1, torch.nn.SyncBatchNorm)
    assert m.K == int(1000 / 1 / 4 * 10) and m.k == 0


def test_bad_transform_layer_raises(gloo1):
    from contrast import resnet
    from contrast.models import PixPro
    with pytest.raises(NotImplementedError):
        PixPro(resnet.resnet18 if False else resnet.resnet50, pixpro_args(pixpro_transform_layer=3))


def test_momentum_schedule(gloo1):
    from contrast import resnet
    from contrast.models import PixPro
    from pixpro_b200._cabi import PixProB200Error
    m = PixPro(resnet.resnet50, pixpro_args())
    m.k, m.K = 3, 10
    mom = 1. - (1. - 0.99) * (math.cos(math.pi * 3 / 10) + 1) / 2.
    assert m._next_momentum() == pytest.approx(mom, abs=1e-15)       # PixPro.py:326
    assert m.k == 4                                                   # PixPro.py:327
    # the update itself is a CUDA kernel: a module kept on the host raises instead of falling back to a torch loop
    with pytest.raises(PixProB200Error):
        m._momentum_update_key_encoder()


def test_unpack_coords_conventions():
    PixProMod = __import__("contrast.models.PixPro", fromlist=["_unpack_coords"])
    cq, ck = torch.zeros(2, 10), torch.ones(2, 10)
    flow, flow_b = torch.zeros(2, 2, 8, 8), torch.ones(2, 2, 8, 8)
    mask = torch.ones(2, 8, 8, dtype=torch.bool)
    assert PixProMod._unpack_coords(cq, ck)[2:] == (None, None, None)
    out = PixProMod._unpack_coords([cq, flow], [ck, flow_b])
    assert out[2] is flow and tuple(out[3]) == (8, 8) and out[4] is None
    size = torch.tensor([720, 1280])
    out = PixProMod._unpack_coords([cq, [flow, size, mask]], [ck, [flow_b, size, None]])
    assert out[0] is cq and out[1] is ck and out[2] is flow and out[3] is size and out[4] is mask
    out = PixProMod._unpack_coords([cq, [flow, size, [mask, "cycle"]]], [ck, [flow_b, size, None]])
    assert out[4] is mask


def test_all_concat_flow_subchains(monkeypatch):
    """use_flow_frames=True enumerates every contiguous sub-chain, shortest first, with the
    bwd slices mirrored (contrast/util.py:111-126)."""
    import contrast.util as util
    calls = []

    def fake_concat(flows, is_norm=False):
        calls.append(flows[:, 0, 0, 0, 0].tolist())
        return torch.zeros(1, 2, 2, 2)

    monkeypatch.setattr(util, "concat_flow", fake_concat)
    n = 4
    fwd = torch.arange(n).float().view(n, 1, 1, 1, 1).expand(n, 1, 2, 2, 2)
    bwd = (100 + torch.arange(n)).float().view(n, 1, 1, 1, 1).expand(n, 1, 2, 2, 2)
    f, b = util.all_concat_flow(fwd, bwd, use_flow_frames=True)
    assert f.shape[0] == n * (n + 1) // 2 == b.shape[0]
    want = []
    for i in range(n):            # the reference's own index arithmetic
        fl = i + 1
        for s in range(n - fl + 1):
            bn = n - s
            want.append(list(range(s, s + fl)))
            want.append([100 + v for v in range(bn - fl, bn)])
    assert calls == [[float(v) for v in w] for w in want]


def test_apply_optical_flow_estimates_links_with_the_flow_model(monkeypatch):
    """The non-file branch (contrast/util.py:76-103, 128-171, 201-204) on CPU with a stand-in flow model: frame pairs go to the
    model forward in time and backward in time (backward links listed from the last frame), `flow_bs` samples at a time, the
    links reach the fused stage in the loader layout [B,n,2,h,w], and the x8 up-sampling is fused only when the model up-samples
    bilinearly (no `update_block.mask`)."""
    import contrast.util as util
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    seen = []

    class FakeRaft(torch.nn.Module):
        def __init__(self, convex):
            super().__init__()
            self.update_block = types.SimpleNamespace(mask=(object() if convex else None))

        def forward(self, a, b, upsample=False, test_mode=True):
            seen.append((float(a[0, 0, 0, 0]), float(b[0, 0, 0, 0]), a.shape[0]))
            low = (b[:, :2, ::8, ::8] - a[:, :2, ::8, ::8]).clone()      # "flow" = frame id difference, 1/8 resolution
            return low, torch.nn.functional.interpolate(low, scale_factor=8.0) * 8

    B = 5
    frames = [torch.full((B, 3, 16, 24), float(t)) + torch.arange(B).view(B, 1, 1, 1) * 10 for t in range(3)]   # frame ids 0, 1, 2
    data = [None] * 7
    data[6] = [torch.tensor([[16, 24]] * B), torch.tensor([[3]] * B)] + frames
    got = {}

    def fake_stage(lo_f, lo_b, flow_up=True, alpha_1=None, alpha_2=None, is_norm=False):
        got.update(lo_f=lo_f, lo_b=lo_b, flow_up=flow_up)
        z = torch.zeros(1)
        return z, z, z, z

    monkeypatch.setattr(util._ops, "flow_stage", fake_stage)
    args = types.SimpleNamespace(alpha1=0.01, alpha2=0.5, use_flow_frames=False, use_flow_file=False, flow_up=True,
                                 flow_cat_norm=False, debug=False, flow_bs=2, verbose=False)
    util.apply_optical_flow(data, FakeRaft(convex=False), args)
    assert got["flow_up"] is True and tuple(got["lo_f"].shape) == (B, 2, 2, 2, 3)
    assert torch.all(got["lo_f"] == 1.0) and torch.all(got["lo_b"] == -1.0)      # every forward link +1 frame, every backward link -1
    # chunks of flow_bs = 2 samples (2, 2, 1); per chunk forward pairs (0,1), (1,2) then backward pairs (2,1), (1,0)
    assert [s[2] for s in seen] == [2] * 4 + [2] * 4 + [1] * 4
    assert [(s[0] % 10, s[1] % 10) for s in seen[:4]] == [(0.0, 1.0), (1.0, 2.0), (2.0, 1.0), (1.0, 0.0)]
    # a model with a learned (convex) up-sampler: its full-resolution prediction is chained, nothing is up-sampled again
    util.apply_optical_flow(data, FakeRaft(convex=True), args)
    assert got["flow_up"] is False and tuple(got["lo_f"].shape) == (B, 2, 2, 16, 24)


def test_regression_loss_debug_tuple_is_rejected():
    PixProMod = __import__("contrast.models.PixPro", fromlist=["regression_loss"])
    with pytest.raises(NotImplementedError):
        PixProMod.regression_loss(torch.zeros(1, 4, 7, 7), torch.zeros(1, 4, 7, 7), (1, 2), (3, 4))


def test_synth_crop_coords_layout():
    from pixpro_b200 import synth
    c = synth.crop_coords(64, seed=3)
    assert c.shape == (64, 10) and c.dtype == torch.float32
    assert torch.all(c[:, 8] == 1280) and torch.all(c[:, 9] == 720)
    flipped = c[:, 0] > c[:, 2]
    assert 0 < int(flipped.sum()) < 64      # both orientations occur
    x0 = torch.minimum(c[:, 0], c[:, 2]) * 1279
    assert torch.allclose(x0, c[:, 4], atol=1e-3)


def test_algorithmic_bytes_match_survey():
    import bench
    # F1 = n*230400 + 14745600 B/sample, F2 = 16.59 MB/sample
    assert bench.kernel_work("chain_up", 1, 1, 7)[0] == 230400 + 14745600
    assert bench.kernel_work("chain_up", 128, 5, 7)[0] == 128 * (5 * 230400 + 14745600)
    assert bench.kernel_work("fb", 1, 1, 7)[0] == 14745600 + 1843200
    # F1+F2 fused and F3+F4 (28 P^2 C flop per sample) for the step-level roofline
    by, fl = bench.step_work(128, 5, 7)
    assert by == 128 * (5 * 230400 + 14745600 + 1843200 + 16 * 256 * 49 * 4) and fl == 128 * 28 * 49 * 49 * 256
    # every tcgen05 contraction of the PPM is one P x P x C product per view: 5 of them = SURVEY's (4+6) P^2 C per view
    ppm = sum(bench.kernel_work(k, 1, 1, 28)[1] for k in ("ppm S (tcgen05)", "ppm Y (tcgen05)", "ppm gS (tcgen05)",
                                                          "ppm gvh (tcgen05)", "ppm gxh (tcgen05)"))
    assert ppm == 2 * (4 + 6) * 784 * 784 * 256
    r = bench.roofline_of("ppm S (tcgen05)", 10, 2.0, 10, 32, 1, 28, {"hbm_gbs": 6545.0, "bf16_tflops_sustained": 1376.3})
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["achieved"] - 32 * 4 * 784 * 784 * 256 / 0.2e-3 / 1e12) < 1e-6
    r = bench.roofline_of("fb", 10, 3.65, 10, 64, 1, 7, {"hbm_gbs": 6545.0})
    assert r["bound"] == "hbm" and abs(r["frac"] - 64 * 16588800 / 0.365e-3 / 1e9 / 6545.0) < 1e-9


def _rank_main(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
    import bench
    a = types.SimpleNamespace(batch=4, n_frames=2, grid=7)
    inp = bench.make_inputs(a.batch, a.n_frames, a.grid, 1234 + rank)   # each rank owns different samples
    ms = bench.max_over_ranks(10.0 + 5.0 * rank, world, torch.device("cpu"))
    fps = bench.aggregate_frames_per_s(a.batch, world, a.n_frames, ms)
    q.put((rank, float(inp["lo_f"].sum()), ms, fps))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing_reduction():
    """world_size-2 gloo: ranks draw disjoint synthetic shards, the step time is the max over
    ranks and the reported throughput is the whole-job aggregate."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, 29547, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, s0, ms0, fps0), (_, s1, ms1, fps1) = res
    assert s0 != s1                                  # different shards
    assert ms0 == ms1 == 15.0                        # max over ranks
    assert fps0 == fps1 == pytest.approx(4 * 2 * 2 / 15e-3)


def test_reference_tree_resolves_behind_the_mirror_package():
    """with this package first on sys.path and the reference tree behind it, mirrored modules
    come from here and unmirrored ones (option, lr_scheduler, non-path names of util) from the reference."""
    import os
    import subprocess
    import sys
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "contrast")):
        pytest.skip("reference tree not present (GPU box)")
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pixpro-with-opticalflow_b200")
    code = (
        "import sys; sys.path[:0] = [%r, %r]\n"
        "import contrast.option as o, contrast.util as u, contrast.lars as l, contrast.lr_scheduler as s\n"
        "from contrast.models import PixPro\n"
        "assert o.__file__.startswith(%r) and s.__file__.startswith(%r)\n"
        "assert u.__file__.startswith(%r) and l.__file__.startswith(%r)\n"
        "assert u.MyHelpFormatter.__module__ == 'contrast._reference_util'\n"
        "print('ok')\n") % (pkg, ref, ref, ref, pkg, pkg)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp", timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-1500:]


def test_sparse_mode_has_no_cpu_fallback_and_keeps_the_list_structure():
    """The lazy stand-ins of the sparse correspondence mode refuse CPU tensors like every other entry (no fallback),
    and regression_loss_pair rejects lazy flows that do not belong together."""
    import types
    import torch
    from pixpro_b200 import ops
    from pixpro_b200._cabi import PixProB200Error
    from contrast import util
    lf = torch.zeros(2, 1, 2, 4, 4)
    with pytest.raises(PixProB200Error):
        ops.LazyFlowPair(lf, lf)
    with pytest.raises(PixProB200Error):
        ops.sparse_corr(lf, lf, torch.zeros(2, 10), None, 7, (32, 32))
    args = types.SimpleNamespace(alpha1=0.01, alpha2=0.5, use_flow_frames=False, use_flow_file=True, flow_up=True,
                                 flow_cat_norm=False, debug=False, flow_sparse=True)
    data = [None] * 7
    data[5] = [None, lf, lf]
    data[6] = [torch.tensor([[32, 32]] * 2), torch.tensor([[2]] * 2)]
    with pytest.raises((PixProB200Error, RuntimeError, AssertionError)):  # `.cuda()` of the mirror, as in the reference
        util.apply_optical_flow(data, None, args)

    class FakePair:  # structure checks run before any device work
        flow_shape = (2, 2, 32, 32)
        lo_fwd = lo_bwd = None
        flow_up, alpha_1, alpha_2 = True, 0.01, 0.5
    a, b = ops.LazyFlow(FakePair(), 0), ops.LazyFlow(FakePair(), 1)
    assert tuple(a.shape) == (2, 2, 32, 32) and a.clone() is a
    assert tuple(ops.LazyMask(FakePair(), 0).shape) == (2, 32, 32)
    q = torch.zeros(2, 4, 7, 7)
    c = torch.zeros(2, 10)
    with pytest.raises(ValueError):  # two different pairs
        ops.regression_loss_pair(q, q, c, c, q, q, c, c, 0.7, flow1=a, flow2=b, size=(32, 32))
    with pytest.raises(ValueError):  # the same direction twice
        p = FakePair()
        ops.regression_loss_pair(q, q, c, c, q, q, c, c, 0.7, flow1=ops.LazyFlow(p, 0), flow2=ops.LazyFlow(p, 0), size=(32, 32))


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm: the oracle port on the host cores) prints ONE JSON line with the
    contract's keys; bounded to a 2-sample step here so that the CPU suite stays fast."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--batch", "2", "--cpu-sample", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_algorithmic_bytes_of_the_other_kernels():
    """Per-kernel figures behind `roofline` for the non-default configurations (DESIGN.md §4 / SURVEY.md §8d)."""
    import bench
    comp, cp4 = 2 * 2 * 720 * 1280 * 4, 256 * 49 * 4
    ab = lambda k, B, n, G=7: bench.kernel_work(k, B, n, G)[0]
    assert ab("chain_dense", 64, 5) == 64 * 6 * comp           # F1': n links in, composites out
    assert ab("fb", 64, 1) == 64 * (comp + 2 * 720 * 1280)      # F2
    # the n = 1 fused route: one direction's up-sampling inside its mask kernel; the three launches move the same bytes plus
    # one extra read of a composite
    lo1 = 2 * 2 * 90 * 160 * 4
    assert ab("chain_up1", 64, 1) == 64 * (lo1 + comp) // 2
    assert ab("fb_up_w", 64, 1) == 64 * (lo1 // 2 + comp + 720 * 1280)
    assert ab("fb1", 64, 1) == 64 * (comp + 720 * 1280)
    assert ab("loss_small", 64, 1, 7) == 64 * 6 * cp4           # F3: q, k in; dq out; both directions
    assert ab("ppm_fwd_small", 64, 1, 7) == 64 * 6 * cp4
    assert ab("ppm_bwd_small", 64, 1, 7) == 64 * 12 * cp4
    assert ab("no such kernel", 64, 1) == 0
    # no kernel the bench can name as dominant is left without a byte or flop figure (`frac: null`)
    for k in ("sparse_corr", "conv1x1 fwd (tcgen05)", "conv1x1 dgrad (tcgen05)", "conv1x1 wgrad (tcgen05)", "conv1x1 bias grad",
              "ppm S (tcgen05)", "ppm Y (tcgen05)", "ppm gS (tcgen05)", "ppm gvh (tcgen05)", "ppm gxh (tcgen05)",
              "loss M=K*pos^T (tcgen05)", "loss_dot", "ppm coldiv", "ppm colnorm", "ppm normbwd"):
        assert ab(k, 64, 1, 14) > 0, k


def test_profiles_index_names_existing_files():
    import re
    txt = open(os.path.join(ROOT, "profiles", "INDEX.md")).read()
    names = set(re.findall(r"`((?:mb/)?r?[\w./-]+\.(?:json|csv|txt|ncu-rep|py|cu))`", txt))
    assert len(names) > 20
    names -= {"bench.py", "_summary.txt"}  # the repo-root script and a suffix mentioned in passing
    missing = [n for n in names if "*" not in n and not os.path.exists(os.path.join(ROOT, "profiles", n))]
    assert not missing, missing


def test_resnet_rejects_architectures_outside_the_scope():
    """unsupported keyword arguments of the reference's ResNet must not be swallowed."""
    from contrast import resnet
    resnet.resnet50(head_type='early_return', low_dim=256)           # sizes a pooled head only: accepted
    for kw in (dict(width=2), dict(deep_stem=True), dict(avg_down=True), dict(groups=32, width_per_group=4), dict(layer4_dilation=2)):
        with pytest.raises(NotImplementedError):
            resnet.resnet50(**kw)
    with pytest.raises(NotImplementedError):
        resnet.resnet50(head_type='pass')                            # the reference pools and flattens there: not provided
    with pytest.raises(TypeError):
        resnet.resnet50(no_such_option=1)


# ---- FastSyncBatchNorm: the collective wiring on two gloo ranks, kernels restated in torch ------------------------------

def _torch_bn_kernels(syncbn):
    """torch restatements of the four C-ABI calls of pixpro_b200.syncbn (same argument / return conventions)."""
    def view(x, N, C):
        return x.reshape(N, C, -1).float()

    def bn_stats(x, layout, N, C, HW, ws):
        v = view(x, N, C)
        m = v.mean((0, 2))
        return torch.cat([m, ((v - m[None, :, None]) ** 2).sum((0, 2)), torch.tensor([float(N * HW)])])

    def bn_apply(x, layout, N, C, HW, stats, nranks, weight, bias, running_mean, running_var, eps, momentum):
        st = stats.reshape(nranks, 2 * C + 1)
        n = st[:, 2 * C]
        cnt = n.sum()
        mean = (n[:, None] * st[:, :C]).sum(0) / cnt
        var = (st[:, C:2 * C] + n[:, None] * (st[:, :C] - mean[None]) ** 2).sum(0) / cnt
        invstd = (var + eps).rsqrt()
        if running_mean is not None:
            running_mean.mul_(1 - momentum).add_(momentum * mean)
            running_var.mul_(1 - momentum).add_(momentum * var * cnt / (cnt - 1))
        y = (view(x, N, C) - mean[None, :, None]) * invstd[None, :, None]
        if weight is not None:
            y = y * weight[None, :, None] + bias[None, :, None]
        return y.reshape(x.shape).to(x.dtype), torch.cat([mean, invstd, cnt.reshape(1)])

    def bn_bwd_stats(dy, x, layout, N, C, HW, save, ws, need_w, need_b):
        g, xm = view(dy, N, C), view(x, N, C) - save[:C][None, :, None]
        s0, s1 = g.sum((0, 2)), (g * xm).sum((0, 2))
        return torch.cat([s0, s1]), (s1 * save[C:2 * C] if need_w else None), (s0.clone() if need_b else None)

    def bn_bwd_apply(dy, x, layout, N, C, HW, save, weight, sums):
        mean, invstd, count = save[:C][None, :, None], save[C:2 * C][None, :, None], save[2 * C]
        mdy = (sums[:C] / count)[None, :, None]
        k = (sums[C:] / count)[None, :, None] * invstd * invstd
        w = 1.0 if weight is None else weight[None, :, None]
        return ((view(dy, N, C) - mdy - (view(x, N, C) - mean) * k) * invstd * w).reshape(x.shape).to(x.dtype)

    syncbn.bn_stats, syncbn.bn_apply, syncbn.bn_bwd_stats, syncbn.bn_bwd_apply = bn_stats, bn_apply, bn_bwd_stats, bn_bwd_apply


def _syncbn_rank_main(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
    from pixpro_b200 import syncbn
    _torch_bn_kernels(syncbn)
    g = torch.Generator().manual_seed(5)
    full = torch.randn(6, 8, 5, 5, generator=g) * 2 + 1      # the global batch; rank r owns samples [3r, 3r+3)
    dfull = torch.randn(6, 8, 5, 5, generator=g)
    w, b = torch.rand(8, generator=g) + 0.5, torch.randn(8, generator=g)
    x = full[3 * rank:3 * rank + 3].clone().requires_grad_(True)
    wl, bl = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    rm, rv = torch.zeros(8), torch.ones(8)
    y = syncbn._SyncBNFunction.apply(x, wl, bl, rm, rv, 1e-5, 0.1, None, None)
    y.backward(dfull[3 * rank:3 * rank + 3])
    # the single-process batch norm over the GLOBAL batch is what synchronised batch norm computes
    xr = full.clone().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(8), torch.ones(8)
    yr = torch.nn.functional.batch_norm(xr, rm2, rv2, wr, br, True, 0.1, 1e-5)
    yr.backward(dfull)
    sl = slice(3 * rank, 3 * rank + 3)
    ok = (torch.allclose(y, yr[sl], atol=1e-5) and torch.allclose(x.grad, xr.grad[sl], atol=1e-5) and
          torch.allclose(rm, rm2, atol=1e-6) and torch.allclose(rv, rv2, atol=1e-5))
    # parameter gradients are LOCAL sums (DDP reduces them): their sum over ranks is the global gradient
    gw, gb = wl.grad.clone(), bl.grad.clone()
    dist.all_reduce(gw); dist.all_reduce(gb)
    ok = ok and torch.allclose(gw, wr.grad, atol=1e-4) and torch.allclose(gb, br.grad, atol=1e-4)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_fast_sync_batch_norm_two_ranks_equal_global_batch_norm():
    """world_size-2 gloo: one all-gather of the [2C+1] statistics forward and one all-reduce of the [2C] sums backward make the two
    ranks' halves equal to batch norm over the whole batch (values, input gradients, running statistics)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_syncbn_rank_main, args=(r, 2, 29551, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]
