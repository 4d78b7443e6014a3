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
