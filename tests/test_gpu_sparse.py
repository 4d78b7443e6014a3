"""GPU parity of the sparse correspondence path (pp_sparse_corr + pp_regression_loss*_warped):
the flow stage evaluated only at the loss's grid centres must be BIT-IDENTICAL to sampling the dense
composites / FB masks — checked against the CPU oracle's dense path (flow_stage + add_optical_flow +
regression_loss, themselves pinned against the reference) at sizes the oracle finishes in seconds, and
against this repo's own dense kernels at BASELINE.json's full batch."""
import types

import numpy as np
import pytest
import torch

from conftest import assert_bits_equal, load_golden, rel_err, unpack_mask

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = "cuda"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def npy(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def synth():
    from pixpro_b200 import synth as s
    return s


def oracle_warp(orc, lo_f, lo_b, coord, G, size, flow_up, a1, a2, direction):
    """Dense oracle: flow stage, then add_optical_flow at the grid centres of `coord`."""
    ff, fb, mf, mb = orc.flow_stage(lo_f, lo_b, flow_up=flow_up, alpha_1=a1, alpha_2=a2)
    flow, mask = (ff, mf) if direction == 0 else (fb, mb)
    B = coord.shape[0]
    z = np.zeros((B, 4, G, G), np.float32)
    o = orc.regression_loss(z, z, coord, coord, 0.7, flow=None, size=size, want_grad=False)  # un-warped centres
    xg, yg = o["cqx"].reshape(B, G, G), o["cqy"].reshape(B, G, G)
    ox, oy, mg = orc.add_optical_flow(flow, xg, yg, size, mask)
    return ox.reshape(B, -1), oy.reshape(B, -1), (None if mg is None else mg.reshape(B, -1)), (ff, fb, mf, mb)


@pytest.mark.parametrize("B,n,h,w,flow_up,use_mask,G,size,mag", [
    (3, 1, 90, 160, True, True, 7, (720, 1280), 1.5),     # BASELINE configs[1] shape
    (2, 5, 90, 160, True, True, 7, (720, 1280), 1.5),     # configs[2]: six frames
    (2, 2, 90, 160, True, True, 14, (720, 1280), 1.5),    # 14x14 grid
    (2, 3, 90, 160, True, False, 7, (720, 1280), 1.5),    # no FB mask
    (2, 2, 90, 160, True, True, 7, (720, 1280), 25.0),    # flows that leave the frame
    (3, 2, 24, 40, True, True, 7, (720, 1280), 1.0),      # flow resolution != image resolution (PixPro.py:76-80)
    (2, 3, 48, 64, False, True, 5, (48, 64), 4.0),        # full-res links, no up-sampling
    (2, 1, 48, 64, False, True, 7, (48, 64), 4.0),
])
def test_sparse_corr_vs_oracle(ops, orc, synth, B, n, h, w, flow_up, use_mask, G, size, mag):
    lf, lb = synth.flow_fields(B, n, h=h, w=w, magnitude=mag, seed=11 * n + G, coarse=(max(2, h // 10), max(2, w // 10)))
    c1 = synth.crop_coords(B, size[1], size[0], seed=5 + n)
    c2 = synth.crop_coords(B, size[1], size[0], seed=6 + n)
    a1, a2 = (0.01, 0.5) if use_mask else (None, None)
    wf, wb = ops.sparse_corr(lf.to(DEV), lb.to(DEV), c1.to(DEV), c2.to(DEV), G, size, flow_up=flow_up, alpha_1=a1, alpha_2=a2)
    for d, (wp, c) in enumerate(((wf, c1), (wb, c2))):
        ox, oy, mg, _ = oracle_warp(orc, lf.numpy(), lb.numpy(), c.numpy(), G, size, flow_up, a1, a2, d)
        assert_bits_equal(npy(wp[0]), ox, f"warped x, direction {d}")
        assert_bits_equal(npy(wp[1]), oy, f"warped y, direction {d}")
        want = np.ones_like(ox) if mg is None else mg.astype(np.float32)
        assert_bits_equal(npy(wp[2]), want, f"mask bit, direction {d}")
        if use_mask and mag < 5 and h == 90:
            assert 0 < want.mean() < 1  # the case exercises both outcomes of the FB test


@pytest.mark.parametrize("tag", ["flow_g7_n1_mask", "flow_g7_n5_mask", "flow_g7_big", "flow_g7_diffsize", "flow_g14_n2_nomask"])
def test_sparse_path_against_reference_golden(ops, tag):
    """The committed fixtures hold what the REFERENCE itself produced (oracle/pin_against_reference.py): warped
    centres and mask bits from its add_optical_flow, loss / pos_num / dq from its regression_loss, for flows its
    apply_optical_flow built from these low-res links.  The sparse path must reproduce them from the links alone."""
    g = load_golden("loss_" + tag)
    size = tuple(int(v) for v in g["size"])
    use_mask = bool(g["use_mask"])
    a1, a2 = (0.01, 0.5) if use_mask else (None, None)
    q = cu(g["q"]).requires_grad_(True)
    B, C, G, _ = q.shape
    P = G * G
    wf, _ = ops.sparse_corr(cu(g["lo_fwd"]), cu(g["lo_bwd"]), cu(g["coord_q"]), None, G, size, alpha_1=a1, alpha_2=a2)
    assert_bits_equal(npy(wf[0]), g["cqx"], "warped centre x")
    assert_bits_equal(npy(wf[1]), g["cqy"], "warped centre y")
    if "mask_grid" in g:
        assert_bits_equal(npy(wf[2]) != 0, g["mask_grid"].astype(bool), "mask_grid")
    pair = ops.LazyFlowPair(cu(g["lo_fwd"]), cu(g["lo_bwd"]), alpha_1=a1, alpha_2=a2)
    loss, pos_num, pos_mean, pos_mask, _ = ops.regression_loss(q, cu(g["k"]), cu(g["coord_q"]), cu(g["coord_k"]),
                                                               float(g["pos_ratio"]), flow=pair.flow[0], size=size,
                                                               mask=pair.mask[0], debug=True)
    loss.backward()
    assert pair._dense is None
    assert_bits_equal(npy(pos_mask), unpack_mask(g["pos_mask"], (B, P, P)), "pos_mask (correspondence indices)")
    assert_bits_equal(npy(pos_num), g["pos_num"], "pos_num")
    assert abs(loss.item() - float(g["loss"])) <= TOL * max(abs(float(g["loss"])), 1e-3)
    assert rel_err(npy(q.grad), g["dq"]) < TOL


def test_sparse_corr_single_direction_and_validation(ops, synth):
    from pixpro_b200._cabi import PixProB200Error
    lf, lb = synth.flow_fields(2, 2, seed=3)
    c1, c2 = synth.crop_coords(2, seed=1).to(DEV), synth.crop_coords(2, seed=2).to(DEV)
    wf, wb = ops.sparse_corr(lf.to(DEV), lb.to(DEV), c1, c2, 7, (720, 1280))
    wf1, none = ops.sparse_corr(lf.to(DEV), lb.to(DEV), c1, None, 7, (720, 1280))
    none2, wb1 = ops.sparse_corr(lf.to(DEV), lb.to(DEV), None, c2, 7, (720, 1280))
    assert none is None and none2 is None and torch.equal(wf, wf1) and torch.equal(wb, wb1)
    with pytest.raises(PixProB200Error):
        ops.sparse_corr(lf.to(DEV), lb.to(DEV), None, None, 7, (720, 1280))
    with pytest.raises(PixProB200Error):
        ops.sparse_corr(lf, lb, c1, c2, 7, (720, 1280))  # CPU tensors: no fallback


@pytest.mark.parametrize("G,n,B", [(7, 2, 4), (14, 1, 3)])
def test_sparse_loss_vs_oracle(ops, orc, synth, G, n, B):
    """Loss, positive mask, counts and gradient of the sparse path against the oracle's dense path."""
    C = 256
    lf, lb = synth.flow_fields(B, n, seed=40 + G)
    cq, ck = synth.crop_coords(B, seed=41), synth.crop_coords(B, seed=42)
    gen = torch.Generator().manual_seed(G)
    q = torch.nn.functional.normalize(torch.randn(B, C, G, G, generator=gen), dim=1)
    k = torch.nn.functional.normalize(torch.randn(B, C, G, G, generator=gen), dim=1)
    off, _, omf, _ = orc.flow_stage(lf.numpy(), lb.numpy())
    o = orc.regression_loss(q.numpy(), k.numpy(), cq.numpy(), ck.numpy(), 0.7, flow=off, size=(720, 1280), mask=omf)
    pair = ops.LazyFlowPair(lf.to(DEV), lb.to(DEV))
    qg = q.to(DEV).requires_grad_(True)
    loss, pos_num, _, pos_mask, centres = ops.regression_loss(qg, k.to(DEV), cq.to(DEV), ck.to(DEV), 0.7, flow=pair.flow[0],
                                                              size=(720, 1280), mask=pair.mask[0], debug=True)
    loss.backward()
    assert pair._dense is None, "the sparse path must not materialise the dense flow stage"
    assert_bits_equal(npy(pos_mask), o["pos_mask"], "pos_mask")
    assert_bits_equal(npy(pos_num), o["pos_num"], "pos_num")
    for i, key in enumerate(["cqx", "cqy", "ckx", "cky"]):
        assert_bits_equal(npy(centres[i]), o[key], key)
    assert abs(loss.item() - o["loss"]) <= TOL * max(abs(o["loss"]), 1e-3)
    assert rel_err(npy(qg.grad), o["dq"]) < TOL


@pytest.mark.parametrize("B,n,G", [(64, 1, 7), (16, 5, 7), (8, 2, 14)])
def test_sparse_pair_is_bit_identical_to_dense_pair_full_size(ops, synth, B, n, G):
    """Size-independent property at the bench sizes: same kernels downstream, so loss / counts / gradients of
    the sparse path equal the dense path's bit for bit, and the dense tensors are recoverable on demand."""
    C = 256
    lf, lb = synth.flow_fields(B, n, seed=70 + n)
    lf, lb = lf.to(DEV), lb.to(DEV)
    f1, f2, k1, k2 = [t.to(DEV) for t in synth.features(B, C, G, seed=71)]
    c1, c2 = synth.crop_coords(B, seed=72).to(DEV), synth.crop_coords(B, seed=73).to(DEV)
    q1 = torch.nn.functional.normalize(f1, dim=1)
    q2 = torch.nn.functional.normalize(f2, dim=1)
    ff, fb, mf, mb = ops.flow_stage(lf, lb)
    qa, qb = q1.clone().requires_grad_(True), q2.clone().requires_grad_(True)
    ld, pnd, pmd = ops.regression_loss_pair(qa, k2, c1, c2, qb, k1, c2, c1, 0.7, flow1=ff, flow2=fb, size=(720, 1280),
                                            mask1=mf, mask2=mb)
    (ld[0] + ld[1]).backward()
    pair = ops.LazyFlowPair(lf, lb)
    qc, qd = q1.clone().requires_grad_(True), q2.clone().requires_grad_(True)
    ls, pns, pms = ops.regression_loss_pair(qc, k2, c1, c2, qd, k1, c2, c1, 0.7, flow1=pair.flow[0], flow2=pair.flow[1],
                                            size=(720, 1280), mask1=pair.mask[0], mask2=pair.mask[1])
    (ls[0] + ls[1]).backward()
    assert pair._dense is None
    assert torch.equal(pnd, pns) and torch.equal(pmd, pms) and torch.equal(ld, ls)
    assert torch.equal(qa.grad, qc.grad) and torch.equal(qb.grad, qd.grad)
    assert pns.sum().item() > 0
    # the intermediate itself: warped centres and mask bits against add_optical_flow on the dense tensors
    wf, wb = ops.sparse_corr(lf, lb, c1, c2, G, (720, 1280))
    z = torch.zeros(B, 256, G, G, device=DEV)
    for wp, c, fl, mk in ((wf, c1, ff, mf), (wb, c2, fb, mb)):
        _, _, _, _, cen = ops.regression_loss(z, z, c, c, 0.7, size=(720, 1280), debug=True)
        ox, oy, mg = ops.add_optical_flow(fl, cen[0].view(B, G, G), cen[1].view(B, G, G), (720, 1280), mk)
        assert torch.equal(wp[0], ox.view(B, -1)) and torch.equal(wp[1], oy.view(B, -1))
        assert torch.equal(wp[2] != 0, mg.view(B, -1))
    # on-demand dense tensors are the flow stage's
    assert torch.equal(pair.flow[0].dense(), ff) and torch.equal(pair.mask[1].dense(), mb)


def test_sparse_mode_of_apply_optical_flow_and_model_loss(ops, synth):
    """contrast.util.apply_optical_flow(args.flow_sparse=True) returns lazy stand-ins with the reference's list
    structure; contrast.models.PixPro.regression_loss consumes them; calc_mask_ratio materialises the masks."""
    from contrast import util
    import importlib
    pp_mod = importlib.import_module("contrast.models.PixPro")  # the module (contrast.models.PixPro is the class)
    B, G, C = 4, 7, 64
    lf, lb = synth.flow_fields(B, 2, seed=3)
    data = [None] * 7
    data[5] = [torch.zeros(B), lf.to(DEV), lb.to(DEV)]
    data[6] = [torch.tensor([[720, 1280]] * B), torch.tensor([[3]] * B)]
    args = types.SimpleNamespace(alpha1=0.01, alpha2=0.5, use_flow_frames=False, use_flow_file=True, flow_up=True,
                                 flow_cat_norm=False, debug=False)
    d1, d2 = util.apply_optical_flow(data, None, args)
    args.flow_sparse = True
    s1, s2 = util.apply_optical_flow(data, None, args)
    assert isinstance(s1[0], ops.LazyFlow) and isinstance(s1[2], ops.LazyMask) and tuple(s1[0].shape) == (B, 2, 720, 1280)
    c1, c2 = synth.crop_coords(B, seed=1).to(DEV), synth.crop_coords(B, seed=2).to(DEV)
    gen = torch.Generator().manual_seed(5)
    q = torch.nn.functional.normalize(torch.randn(B, C, G, G, generator=gen), dim=1).to(DEV)
    k = torch.nn.functional.normalize(torch.randn(B, C, G, G, generator=gen), dim=1).to(DEV)
    for dense, lazy, cq, ck in ((d1, s1, c1, c2), (d2, s2, c2, c1)):
        l_d, (pn_d, pm_d) = pp_mod.regression_loss(q, k, [cq, dense], [ck, dense], 0.7)
        l_s, (pn_s, pm_s) = pp_mod.regression_loss(q, k, [cq, lazy], [ck, lazy], 0.7)
        assert torch.equal(l_d, l_s) and torch.equal(pn_d, pn_s) and torch.equal(pm_d, pm_s)
    assert s1[0].pair._dense is None
    assert torch.equal(util.calc_mask_ratio(s1[2].clone()), util.calc_mask_ratio(d1[2]))
    assert torch.equal(util.calc_mask_ratio(s2[2]), util.calc_mask_ratio(d2[2]))


def test_host_pixel_step_sparse_matches_dense():
    from pixpro_b200 import synth
    from pixpro_b200.host_step import HostPixelStep
    B, C, G = 8, 256, 7
    lf, lb = synth.flow_fields(B, 2, seed=5)
    f1, f2, k1, k2 = synth.features(B, C, G, seed=6)
    c1, c2 = synth.crop_coords(B, seed=7), synth.crop_coords(B, seed=8)
    gen = torch.Generator().manual_seed(9)
    w = (torch.randn(C, C, 1, 1, generator=gen) / 16).to(DEV)
    bias = torch.zeros(C, device=DEV)
    host = {k: v.pin_memory() for k, v in dict(lo_f=lf, lo_b=lb, feat1=f1, feat2=f2, k1=k1, k2=k2, c1=c1, c2=c2).items()}
    outs = []
    for sparse in (False, True):
        step = HostPixelStep(DEV, B, C, G, use_graph=True, sparse=sparse)
        for _ in range(4):
            out, _ = step(host, w, bias)
        outs.append({k: v.clone() for k, v in out.items()})
    assert torch.equal(outs[0]["pos_num"], outs[1]["pos_num"])
    assert torch.equal(outs[0]["loss"], outs[1]["loss"])
    assert torch.equal(outs[0]["d_feat"], outs[1]["d_feat"])


def test_sparse_equals_dense_in_rcp_division_mode(ops, synth):
    """PIXPRO_B200_DIV=rcp (tensor / scalar as torch's CUDA kernels round it): the sparse kernel follows the same mode."""
    B, n, G, C = 6, 3, 7, 64
    lf, lb = synth.flow_fields(B, n, seed=90)
    lf, lb = lf.to(DEV), lb.to(DEV)
    c1, c2 = synth.crop_coords(B, seed=91).to(DEV), synth.crop_coords(B, seed=92).to(DEV)
    gen = torch.Generator().manual_seed(93)
    q = torch.nn.functional.normalize(torch.randn(B, C, G, G, generator=gen), dim=1).to(DEV)
    k = torch.nn.functional.normalize(torch.randn(B, C, G, G, generator=gen), dim=1).to(DEV)
    ops.set_div_mode("rcp")
    try:
        ff, fb, mf, mb = ops.flow_stage(lf, lb)
        ld, pnd, _ = ops.regression_loss_pair(q, k, c1, c2, q, k, c2, c1, 0.7, flow1=ff, flow2=fb, size=(720, 1280), mask1=mf, mask2=mb)
        pair = ops.LazyFlowPair(lf, lb)
        ls, pns, _ = ops.regression_loss_pair(q, k, c1, c2, q, k, c2, c1, 0.7, flow1=pair.flow[0], flow2=pair.flow[1],
                                              size=(720, 1280), mask1=pair.mask[0], mask2=pair.mask[1])
        assert torch.equal(pnd, pns) and torch.equal(ld, ls) and pns.sum().item() > 0
    finally:
        ops.set_div_mode("ieee")


@pytest.mark.parametrize("B,n,G", [(64, 1, 7), (16, 5, 7), (8, 2, 14)])
def test_sparse_corr_full_size_vs_oracle_restatement(ops, orc, synth, B, n, G):
    """At the bench sizes the dense oracle is too slow for a unit test, but the oracle's own sparse restatement
    (orc_sparse_corr, pinned against the reference's dense route on CPU) is not: bit-exact at full batch."""
    lf, lb = synth.flow_fields(B, n, seed=120 + n)
    c1, c2 = synth.crop_coords(B, seed=121), synth.crop_coords(B, seed=122)
    wf, wb = ops.sparse_corr(lf.to(DEV), lb.to(DEV), c1.to(DEV), c2.to(DEV), G, (720, 1280))
    of, ob = orc.sparse_corr(lf.numpy(), lb.numpy(), c1.numpy(), c2.numpy(), G, (720, 1280))
    assert_bits_equal(npy(wf), of, "forward direction")
    assert_bits_equal(npy(wb), ob, "backward direction")
    assert 0.2 < of[2].mean() < 1.0
