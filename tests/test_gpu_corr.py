"""GPU: RAFT correlation volume / pyramid / lookup kernels (csrc/pp_corr.cu) through the C ABI
against the goldens of the reference's torch CorrBlock (contrast/flow/corr.py:12-60) and against the oracle.
Volume: 1e-5 relative to the result's scale (fp32 contraction, the path's float bar); pooling and lookup: bit-exact."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import assert_bits_equal, load_golden, rel_err
from test_oracle_golden import CORR_TAGS

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-5


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def npy(t):
    return t.detach().cpu().numpy()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("tag", CORR_TAGS)
def test_corr_block_golden(tag):
    """The drop-in class, used exactly as contrast/flow/raft.py:120-141 uses the reference's."""
    from contrast.flow.corr import CorrBlock
    g = load_golden("corr_" + tag)
    L, r = int(g["num_levels"]), int(g["radius"])
    B, D, h, w = g["fmap1"].shape
    blk = CorrBlock(cu(g["fmap1"]), cu(g["fmap2"]), num_levels=L, radius=r)
    assert len(blk.corr_pyramid) == L
    for l, p in enumerate(blk.corr_pyramid):
        assert tuple(p.shape) == (B * h * w, 1, h >> l, w >> l)
    assert rel_err(npy(blk.corr_pyramid[0]), g["level0"]) < TOL
    out = npy(blk(cu(g["coords"])))
    assert out.shape == g["out"].shape and out.dtype == np.float32
    # the lookup interpolates the kernel's own volume (1e-5 from the reference's): tolerance of the volume
    assert rel_err(out, g["out"]) < TOL


@pytest.mark.parametrize("tag", CORR_TAGS)
def test_corr_pool_and_lookup_bit_exact_from_reference_volume(tag):
    """Pooling and lookup fed with the REFERENCE's level-0 volume are bit-exact against the reference's outputs."""
    from pixpro_b200 import ops
    g = load_golden("corr_" + tag)
    L, r = int(g["num_levels"]), int(g["radius"])
    pyr = [cu(g["level0"])]
    for l in range(1, L):
        pyr.append(ops.corr_pool(pyr[-1]))
        assert sha(npy(pyr[-1])) == str(g[f"level{l}_sha"]), f"pyramid level {l}"
    assert_bits_equal(npy(ops.corr_lookup(pyr, cu(g["coords"]), r)), g["out"], "lookup")


@pytest.mark.parametrize("B,D,h,w,L,r", [(2, 128, 46, 62, 4, 3),     # RAFT-small on a 368x496 frame (1/8 resolution)
                                         (1, 256, 16, 16, 4, 4),     # RAFT-basic feature width and radius
                                         (3, 24, 5, 9, 2, 2),        # plane below 128 positions: CUDA-core contraction
                                         (1, 128, 11, 13, 3, 3)])    # odd sizes: pooling drops the last row / column
def test_corr_vs_oracle(orc, B, D, h, w, L, r):
    from pixpro_b200 import ops
    g = torch.Generator().manual_seed(B * 100 + h)
    f1 = torch.randn(B, D, h, w, generator=g)
    f2 = torch.randn(B, D, h, w, generator=g)
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    coords = torch.stack([xs, ys]).float()[None].repeat(B, 1, 1, 1) + 4.0 * torch.randn(B, 2, h, w, generator=g)
    vol = ops.corr_volume(f1.to(DEV), f2.to(DEV))
    assert rel_err(npy(vol), orc.corr_volume(f1.numpy(), f2.numpy())) < TOL
    pyr_g, pyr_o = [vol.view(B * h * w, 1, h, w)], [npy(vol).reshape(B * h * w, 1, h, w)]
    for l in range(1, L):
        pyr_g.append(ops.corr_pool(pyr_g[-1]))
        pyr_o.append(orc.corr_pool(pyr_o[-1]))
        assert_bits_equal(npy(pyr_g[-1]), pyr_o[-1], f"pyramid level {l}")
    assert_bits_equal(npy(ops.corr_lookup(pyr_g, coords.to(DEV), r)), orc.corr_lookup(pyr_o, coords.numpy(), r), "lookup")


def test_corr_empty_batch_and_errors():
    from pixpro_b200 import ops
    from pixpro_b200._cabi import PixProB200Error
    # a pyramid level with a single row or column: the reference divides by (size - 1) = 0 there (utils.py:68-69) and
    # samples NaN coordinates; the kernel refuses instead of reproducing that
    f = torch.randn(1, 16, 6, 10, device=DEV)
    pyr = [ops.corr_volume(f, f).view(60, 1, 6, 10)]
    for _ in range(2):
        pyr.append(ops.corr_pool(pyr[-1]))
    with pytest.raises(PixProB200Error):
        ops.corr_lookup(pyr, torch.zeros(1, 2, 6, 10, device=DEV), 2)
    z = torch.zeros(0, 16, 8, 8, device=DEV)
    assert tuple(ops.corr_volume(z, z).shape) == (0, 64, 64)
    with pytest.raises((PixProB200Error, AssertionError)):
        ops.corr_volume(torch.zeros(1, 16, 8, 8), torch.zeros(1, 16, 8, 8))  # CPU tensors: no fallback


@pytest.mark.parametrize("tag", ["small", "basic"])
def test_raft_estimator_matches_the_reference(tag):
    """The drop-in contrast.flow.RAFT (cuDNN convolutions + this package's correlation volume / pyramid / lookup / x8
    up-sampling kernels) against the reference's RAFT run on CPU with the same name-seeded weights.  Tolerance: 2e-3 of the
    flow's scale — cuDNN and CPU convolutions differ in summation order and the GRU iterations feed back on themselves."""
    import types
    from contrast.flow import RAFT
    from pixpro_b200 import synth
    g = load_golden("raft_" + tag)
    m = RAFT(types.SimpleNamespace(small=bool(g["small"]), mixed_precision=False))
    synth.seeded_init_(m, int(g["seed"]))
    m = m.to(DEV).eval()
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            low, up = m(cu(g["image1"].astype(np.float32)), cu(g["image2"].astype(np.float32)), iters=int(g["iters"]), upsample=False,
                        test_mode=True)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    assert tuple(low.shape) == g["flow_low"].shape and tuple(up.shape) == (low.shape[0], 2, 8 * low.shape[2], 8 * low.shape[3])
    scale = float(np.abs(g["flow_low"]).max())
    assert np.abs(npy(low) - g["flow_low"]).max() < 2e-3 * scale
    assert abs(float(up.abs().max()) - float(g["flow_up_absmax"])) < 2e-3 * float(g["flow_up_absmax"])


def test_apply_optical_flow_estimates_links_with_the_raft_model():
    """The non-file path of apply_optical_flow (contrast/util.py:201-204, 76-103, 128-171): links come from the flow model
    (here a name-seeded RAFT-small on three 128x160 frames), `flow_bs` chunks do not change them, and the stage that
    follows is the precomputed-links stage applied to those links."""
    import types
    from contrast import util
    from contrast.flow import RAFT
    from pixpro_b200 import ops, synth
    m = RAFT(types.SimpleNamespace(small=True, mixed_precision=False))
    synth.seeded_init_(m, 300)
    with torch.no_grad():  # a randomly initialised RAFT iterated 12 times is chaotic: damp its updates so that runs are comparable
        m.update_block.flow_head.conv2.weight.mul_(0.02)
        m.update_block.flow_head.conv2.bias.mul_(0.02)
    m = m.to(DEV).eval()
    g = torch.Generator().manual_seed(3)
    B = 3
    frames = [torch.rand(B, 3, 128, 160, generator=g) * 255.0 for _ in range(3)]
    size = torch.tensor([[128, 160]] * B)
    data = [None] * 5 + [None, [size, torch.tensor([[3]] * B)] + frames]
    args = types.SimpleNamespace(alpha1=0.01, alpha2=0.5, use_flow_frames=False, use_flow_file=False, flow_up=True, flow_cat_norm=False,
                                 debug=False, flow_bs=B, verbose=False)
    (ff, sz, mf), (fb, _, mb) = util.apply_optical_flow(data, m, args)
    assert tuple(ff.shape) == (B, 2, 128, 160) and tuple(mf.shape) == (B, 128, 160) and torch.equal(sz, size[0])
    # the same links estimated in one chunk, then the fused stage
    with torch.no_grad():
        fr = [f.to(DEV) for f in frames]
        lo_f = torch.stack([m(a, b, upsample=False, test_mode=True)[0] for a, b in zip(fr[:-1], fr[1:])], dim=1)
        lo_b = torch.stack([m(a, b, upsample=False, test_mode=True)[0] for a, b in zip(fr[::-1][:-1], fr[::-1][1:])], dim=1)
    want = ops.flow_stage(lo_f, lo_b, flow_up=True, alpha_1=0.01, alpha_2=0.5)
    scale = want[0].abs().max().item()
    assert (ff - want[0]).abs().max().item() < 1e-4 * scale and (fb - want[1]).abs().max().item() < 1e-4 * scale
    assert (mf != want[2]).float().mean().item() < 1e-3
    cf, cb = util.mem_reduce_calc_optical_flow(frames, m, args)
    assert tuple(cf.shape) == (1, B, 2, 128, 160)
    assert (cf[0] - ff).abs().max().item() < 1e-4 * scale and (cb[0] - fb).abs().max().item() < 1e-4 * scale
    # estimating flow_bs = 2 samples at a time: same layout; values agree up to what a randomly initialised, iterated
    # network makes of cuDNN choosing another algorithm for another batch size
    args.flow_bs = 2
    (ff2, _, mf2), _ = util.apply_optical_flow(data, m, args)
    assert tuple(ff2.shape) == tuple(ff.shape) and tuple(mf2.shape) == tuple(mf.shape)
    assert (ff2 - ff).abs().max().item() < 1e-2 * scale
