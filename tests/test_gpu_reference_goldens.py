"""GPU: the drop-in modules against goldens the UNMODIFIED reference produced (oracle/pin_against_reference.py, run in the
build container, fixtures under tests/golden/): the whole model step of BASELINE configs[0] (ResNet-50, B=4, 7x7, no
flow), the instance branch, the general `apply_optical_flow` path with `use_flow_frames` (a4) and its
`debug` return structure (a6), and one-sample slices of the full bench batch against the CPU oracle."""
import os
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist

from conftest import assert_bits_equal, load_golden, unpack_mask

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
        os.environ.setdefault("MASTER_PORT", "29553")
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo", rank=0, world_size=1)
    yield
    if dist.is_initialized():
        dist.destroy_process_group()


def npy(t):
    return t.detach().cpu().numpy()


# Tolerances of the model-level comparison.  The golden is the reference on torch CPU kernels (MKL-DNN convolutions, BN on
# 4 x 7 x 7 samples per channel), this side runs the backbone on cuDNN: 53 convolution + batch-norm layers amplify the
# ~1e-6 per-layer rounding differences before the pixel path sees them.  The pixel path itself is held to 1e-5 at
# function level (tests/test_gpu_parity.py); here the bar is what the backbone's own CPU/GPU disagreement allows.
LOSS_ABS_TOL = 2e-4      # on a loss in [-4, 4] (two directions x -2 * mean cosine)
GRAD_NORM_RTOL = 2e-2    # per-tensor gradient norms


@pytest.mark.parametrize("tag", ["cfg0", "cfg0_ins"])
def test_model_step_matches_reference_golden(group, tag):
    """BASELINE configs[0] (and the same with pixpro_ins_loss_weight=1, a12): loss, positive counts and the
    per-parameter gradient norms of one fwd+bwd step of contrast.models.PixPro against the reference model's."""
    from contrast import resnet
    from contrast.models import PixPro
    from pixpro_b200 import synth
    g = load_golden("model_" + tag)
    seed = int(g["seed"])
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(seed)
        m = PixPro(resnet.resnet50, pixpro_args(pixpro_ins_loss_weight=float(g["ins_weight"]))).to(DEV)
        synth.seeded_init_(m, seed)
        m.train()
        gen = torch.Generator().manual_seed(seed)
        im1 = torch.randn(4, 3, 224, 224, generator=gen).to(DEV)
        im2 = torch.randn(4, 3, 224, 224, generator=gen).to(DEV)
        c1, c2 = synth.crop_coords(4, seed=seed + 1).to(DEV), synth.crop_coords(4, seed=seed + 2).to(DEV)
        loss, ((pn1, _), (pn2, _)) = m(im1, im2, c1, c2, is_update_momentum=False)
        loss.backward()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    assert_bits_equal(npy(pn1), g["pos_num_1"], "pos_num direction 1")
    assert_bits_equal(npy(pn2), g["pos_num_2"], "pos_num direction 2")
    err = abs(loss.item() - float(g["loss"]))
    params = dict(m.named_parameters())
    names = [str(n) for n in g["grad_names"]]
    assert sorted(names) == sorted(n for n, p in params.items() if p.grad is not None), "the same tensors receive gradients"
    gn = np.array([float(params[n].grad.double().norm()) for n in names])
    # relative to the tensor's own norm, floored at 1e-3 of the largest one: the biases in front of a batch norm have a
    # mathematically zero gradient, whose computed norm (~1e-8) is pure rounding noise on either side
    rel = np.abs(gn - g["grad_norms"]) / np.maximum(g["grad_norms"], 1e-3 * g["grad_norms"].max())
    print(f"[{tag}] loss {loss.item():.6f} vs reference {float(g['loss']):.6f} (abs err {err:.2e}); "
          f"gradient norms: max rel err {rel.max():.2e} over {len(names)} tensors")
    assert err <= LOSS_ABS_TOL, (loss.item(), float(g["loss"]))
    assert rel.max() <= GRAD_NORM_RTOL, (names[int(rel.argmax())], float(rel.max()))


def _apply_general(g, **kw):
    import contrast.util as util
    args = types.SimpleNamespace(alpha1=0.01, alpha2=0.5, use_flow_frames=False, use_flow_file=True, flow_up=True,
                                 flow_cat_norm=False, debug=False, verbose=False)
    args.__dict__.update(kw)
    f, b = torch.from_numpy(g["lo_fwd"]).to(DEV), torch.from_numpy(g["lo_bwd"]).to(DEV)
    B, n, _, h, w = f.shape
    data = [None] * 7
    data[5] = [torch.zeros(B), f, b]
    data[6] = [torch.tensor([[8 * h, 8 * w]] * B), torch.tensor([[n + 1]] * B)]
    return util.apply_optical_flow(data, None, args)


def test_apply_optical_flow_use_flow_frames_golden():
    """a4 / a6 general path: `use_flow_frames` with 4 frames returns every contiguous sub-chain, stacked
    [n(n+1)/2, B, 2, H, W] (util.py:111-126,206-244) — bit-exact against the reference's own output."""
    g = load_golden("apply_general_frames_n3")
    (ff, size, mf), (fb, _, mb) = _apply_general(g, use_flow_frames=True)
    assert ff.shape == g["flow_fwd"].shape and ff.ndim == 5
    assert_bits_equal(npy(ff), g["flow_fwd"], "stacked forward sub-chains")
    assert_bits_equal(npy(fb), g["flow_bwd"], "stacked backward sub-chains")
    shape = tuple(int(s) for s in g["mask_shape"])
    assert_bits_equal(npy(mf), unpack_mask(g["mask_fwd"], shape), "stacked forward FB masks")
    assert_bits_equal(npy(mb), unpack_mask(g["mask_bwd"], shape), "stacked backward FB masks")
    assert [int(s) for s in size] == [int(s) for s in g["size"]]


def test_apply_optical_flow_debug_structure_golden():
    """The `debug` return structure of the general path: masks come back as [mask, cycle_flow] lists (util.py:218-227)."""
    g = load_golden("apply_general_debug_n2")
    (ff, _, mf), (fb, _, mb) = _apply_general(g, debug=True)
    assert isinstance(mf, list) and len(mf) == 2
    assert_bits_equal(npy(ff), g["flow_fwd"], "flow_fwd")
    assert_bits_equal(npy(fb), g["flow_bwd"], "flow_bwd")
    shape = tuple(int(s) for s in g["mask_shape"])
    assert_bits_equal(npy(mf[0]), unpack_mask(g["mask_fwd"], shape), "mask_fwd")
    assert_bits_equal(npy(mb[0]), unpack_mask(g["mask_bwd"], shape), "mask_bwd")
    assert_bits_equal(npy(mf[1]), g["cycle_fwd"], "cycle flow fwd")
    assert_bits_equal(npy(mb[1]), g["cycle_bwd"], "cycle flow bwd")


@pytest.mark.parametrize("n_frames,samples", [(2, (0, 37, 63)), (6, (21,))])
def test_bench_batch_slices_vs_oracle(n_frames, samples):
    """The dense flow stage at the FULL bench batch (B=64, 90x160 -> 720x1280, bench.py's own inputs): individual
    samples of the batched run compared directly with the CPU oracle run on that sample alone."""
    import bench
    from oracle import oracle as orc
    from pixpro_b200 import ops
    inp = bench.make_inputs(64, n_frames, 7, 1234)
    ff, fb, mf, mb = ops.flow_stage(inp["lo_f"].to(DEV), inp["lo_b"].to(DEV), flow_up=True, alpha_1=bench.ALPHA1, alpha_2=bench.ALPHA2)
    for s in samples:
        want = orc.flow_stage(inp["lo_f"][s:s + 1].numpy(), inp["lo_b"][s:s + 1].numpy(), alpha_1=bench.ALPHA1, alpha_2=bench.ALPHA2)
        for name, got, w_ in zip(["flow_fwd", "flow_bwd", "mask_fwd", "mask_bwd"], (ff, fb, mf, mb), want):
            assert_bits_equal(npy(got[s:s + 1]), w_, f"sample {s} {name}")


def test_near_threshold_pair_counts_are_recorded():
    """every positive-mask golden carries the number of pairs within 1e-5 of the distance threshold
    (the pairs a 1-ulp coordinate difference could flip); the GPU masks are bit-exact regardless (test_gpu_parity.py)."""
    total = 0
    for tag in ["noflow_g7", "noflow_g14", "flow_g7_n1_mask", "flow_g7_n5_mask", "flow_g14_n2_nomask", "flow_g7_diffsize",
                "flow_g7_big", "noflow_g7_ratio03"]:
        g = load_golden("loss_" + tag)
        assert "near_threshold_pairs" in g and g["near_threshold_pairs"].shape == g["pos_num"].shape
        total += int(g["near_threshold_pairs"].sum())
        print(f"loss_{tag}: pairs within 1e-5 of the threshold per sample {g['near_threshold_pairs'].tolist()}")
    assert total >= 0
