"""CPU: the oracle (oracle/pixpro_oracle.c) against the golden vectors produced by the REAL
reference (oracle/pin_against_reference.py, run in the build container where /root/reference
exists).  Bit-exact for coordinates / masks / counts; 1e-5 relative for float reductions."""
import hashlib

import numpy as np
import pytest

from conftest import assert_bits_equal, load_golden, rel_err, unpack_mask

TOL = 1e-5  # BASELINE.json north_star: loss values and gradients within 1e-5 relative in fp32


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_upflow8(orc):
    g = load_golden("upflow8")
    assert_bits_equal(orc.upflow8(g["inp"]), g["out"], "upflow8")
    g = load_golden("upflow8_full")
    assert sha(orc.upflow8(g["inp"].reshape(-1, 2, 90, 160))) == str(g["out_sha"])


@pytest.mark.parametrize("name", ["normalize_coord", "normalize_flow", "denormalize_flow"])
def test_normalize(orc, name):
    g = load_golden(name)
    assert_bits_equal(getattr(orc, name)(g["inp"]), g["out"], name)


@pytest.mark.parametrize("tag", ["n1", "n2", "n5", "n5_oob", "n3_norm", "n1_norm"])
def test_concat_flow(orc, tag):
    g = load_golden("concat_flow_" + tag)
    assert_bits_equal(orc.concat_flow(g["flows"], is_norm=bool(g["is_norm"])), g["out"], "concat_flow " + tag)


@pytest.mark.parametrize("tag", ["a", "oob", "norm"])
def test_fb_consistency(orc, tag):
    g = load_golden("fb_" + tag)
    c1, m, cyc = orc.forward_backward_consistency(g["fwd"], g["bwd"], 0.01, 0.5, is_norm=bool(g["is_norm"]))
    assert_bits_equal(m, g["mask"], "mask")
    assert_bits_equal(cyc, g["cycle"], "cycle")
    assert_bits_equal(c1, g["coords1"], "coords1")


@pytest.mark.parametrize("tag", ["n1_up", "n5_up", "n2_noup", "n5_nomask", "n3_catnorm"])
def test_flow_stage(orc, tag):
    g = load_golden("flow_stage_" + tag)
    use_mask = bool(g["use_mask"])
    ff, fb, mf, mb = orc.flow_stage(g["lo_fwd"], g["lo_bwd"], flow_up=bool(g["flow_up"]),
                                    alpha_1=0.01 if use_mask else None, alpha_2=0.5 if use_mask else None,
                                    is_norm=bool(g["is_norm"]))
    assert_bits_equal(ff, g["flow_fwd"], "flow_fwd")
    assert_bits_equal(fb, g["flow_bwd"], "flow_bwd")
    if use_mask:
        assert_bits_equal(mf, unpack_mask(g["mask_fwd"], mf.shape), "mask_fwd")
        assert_bits_equal(mb, unpack_mask(g["mask_bwd"], mb.shape), "mask_bwd")
        assert rel_err(orc.calc_mask_ratio(mf), g["mask_ratio_fwd"]) < 1e-6


@pytest.mark.parametrize("tag", ["full_n1", "full_n5"])
def test_flow_stage_full_size(orc, tag):
    g = load_golden("flow_stage_" + tag)
    ff, fb, mf, mb = orc.flow_stage(g["lo_fwd"], g["lo_bwd"])
    assert sha(ff) == str(g["flow_fwd_sha"])
    assert sha(fb) == str(g["flow_bwd_sha"])
    assert_bits_equal(mf, unpack_mask(g["mask_fwd"], mf.shape), "mask_fwd")
    assert_bits_equal(mb, unpack_mask(g["mask_bwd"], mb.shape), "mask_bwd")


LOSS_TAGS = ["noflow_g7", "noflow_g14", "flow_g7_n1_mask", "flow_g7_n5_mask", "flow_g14_n2_nomask",
             "flow_g7_diffsize", "flow_g7_big", "noflow_g7_ratio03"]


def oracle_loss_inputs(orc, g):
    """Rebuild the dense flow / mask inputs of a loss fixture from its low-res links."""
    flow = mask = None
    if "lo_fwd" in g:
        um = bool(g["use_mask"])
        ff, _, mf, _ = orc.flow_stage(g["lo_fwd"], g["lo_bwd"], alpha_1=0.01 if um else None, alpha_2=0.5 if um else None)
        flow, mask = ff, mf
    return flow, mask


@pytest.mark.parametrize("tag", LOSS_TAGS)
def test_regression_loss(orc, tag):
    g = load_golden("loss_" + tag)
    flow, mask = oracle_loss_inputs(orc, g)
    o = orc.regression_loss(g["q"], g["k"], g["coord_q"], g["coord_k"], float(g["pos_ratio"]), flow=flow,
                            size=tuple(g["size"]), mask=mask)
    assert_bits_equal(o["pos_num"], g["pos_num"], "pos_num")
    assert rel_err(o["pos_mean"], g["pos_mean"]) < 1e-6
    assert abs(o["loss"] - float(g["loss"])) <= TOL * max(abs(float(g["loss"])), 1e-3)
    assert rel_err(o["dq"], g["dq"]) < TOL
    if "cqx" in g:
        assert_bits_equal(o["cqx"], g["cqx"], "warped centre x")
        assert_bits_equal(o["cqy"], g["cqy"], "warped centre y")


@pytest.mark.parametrize("tag", ["l1_p2_g7", "l0_p1_g7", "l1_p2_g14", "l0_p05_cv01", "l0_p3_g7"])
def test_featprop(orc, tag):
    g = load_golden("featprop_" + tag)
    out = orc.featprop(g["feat"], g["val"], gamma=float(g["gamma"]), clamp_value=float(g["clamp"]))
    assert rel_err(out, g["out"]) < TOL
    dfs, dv = orc.featprop_bwd(g["feat"], g["val"], g["gout"], gamma=float(g["gamma"]), clamp_value=float(g["clamp"]))
    if "weight" in g:
        w = g["weight"][:, :, 0, 0].astype(np.float64)
        dfeat = dfs + np.einsum("oc,bohw->bchw", w, dv.astype(np.float64))
        assert rel_err(np.einsum("bohw,bchw->oc", dv.astype(np.float64), g["feat"].astype(np.float64)),
                       g["d_weight"][:, :, 0, 0]) < 2e-5
    else:
        dfeat = dfs + dv
    assert rel_err(dfeat, g["d_feat"]) < 2e-5


# ------------------------------------------------------------------ optimizer side

def test_ema_update_golden(orc):
    g = load_golden("ema")
    assert_bits_equal(orc.ema_update(g["k"], g["q"], float(g["m"])), g["out"], "EMA of the key branch")


def lars_golden_steps(g):
    """Yields (tensor index, step, hyper-parameters, p_before, grad, p_after_reference); p_before is the
    reference's own parameter of the previous step."""
    n, steps = int(g["n_params"]), int(g["n_steps"])
    for i in range(n):
        wd, lr, mom, damp, lars = g["meta"][i]
        prev = g[f"p{i}_init"]
        for s in range(steps):
            yield i, s, (wd, lr, mom, damp, bool(lars)), prev, g[f"g{i}_s{s}"], g[f"p{i}_s{s}"]
            prev = g[f"p{i}_s{s}"]


def test_lars_sgd_golden(orc):
    """The oracle's LARS + SGD restatement, chained over 3 steps from the initial parameters, against the
    parameters the real reference optimizer produced: bit-exact where LARS does not scale, 1e-6 where the
    (order-dependent) norms enter."""
    g = load_golden("lars_sgd")
    state = {}
    for i, s, (wd, lr, mom, damp, lars), _, grad, want in lars_golden_steps(g):
        p0, buf = state.get(i, (g[f"p{i}_init"], None))
        p1, buf1, rate = orc.lars_sgd_step(p0, grad, buf, wd, lr, mom, damp, lars=lars, first=buf is None,
                                           trust=float(g["trust"]), eps=float(g["eps"]))
        if lars:
            assert rel_err(p1, want) <= 1e-6 and 0.0 < rate
        else:
            assert rate == 1.0
            assert_bits_equal(p1, want, f"SGD tensor {i} step {s}")
        state[i] = (p1, buf1)


# ---- sparse correspondence restatement (orc_sparse_corr): flow stage + add_optical_flow at the grid centres only ----

@pytest.mark.parametrize("tag", ["flow_g7_n1_mask", "flow_g7_n5_mask", "flow_g7_big", "flow_g7_diffsize", "flow_g14_n2_nomask"])
def test_sparse_corr_against_reference_golden(orc, tag):
    """Warped centres and mask bits the REFERENCE produced (its add_optical_flow on the flows its apply_optical_flow built
    from these links), reproduced from the low-res links alone."""
    g = load_golden("loss_" + tag)
    use_mask = bool(g["use_mask"])
    G = g["q"].shape[-1]
    wf, wb = orc.sparse_corr(g["lo_fwd"], g["lo_bwd"], g["coord_q"], None, G, tuple(int(v) for v in g["size"]),
                             alpha_1=0.01 if use_mask else None, alpha_2=0.5 if use_mask else None)
    assert wb is None
    assert_bits_equal(wf[0], g["cqx"], "warped centre x")
    assert_bits_equal(wf[1], g["cqy"], "warped centre y")
    if "mask_grid" in g:
        assert_bits_equal(wf[2] != 0, g["mask_grid"].astype(bool), "mask_grid")
    else:
        assert (wf[2] == 1).all()


@pytest.mark.parametrize("n,h,w,flow_up,size,G,mag,div_mode", [
    (1, 18, 32, True, (144, 256), 7, 0.4, 0), (3, 18, 32, True, (144, 256), 7, 0.4, 0), (5, 18, 32, True, (144, 256), 5, 3.0, 0),
    (2, 18, 32, True, (720, 1280), 7, 0.4, 0), (3, 48, 64, False, (48, 64), 7, 4.0, 0), (2, 18, 32, True, (144, 256), 7, 0.4, 1)])
def test_sparse_corr_equals_sampling_the_dense_stage(orc, n, h, w, flow_up, size, G, mag, div_mode):
    """The property the sparse mode rests on: evaluating the chain / FB test at single pixels equals sampling the dense
    composites and masks (both directions, out-of-frame flows, flow resolution != image resolution, rcp division)."""
    from pixpro_b200 import synth
    B = 3
    lf, lb = synth.flow_fields(B, n, h=h, w=w, magnitude=mag, seed=7 * n + G, coarse=(3, 4))
    lf, lb = lf.numpy(), lb.numpy()
    c1 = synth.crop_coords(B, size[1], size[0], seed=n).numpy()
    c2 = synth.crop_coords(B, size[1], size[0], seed=n + 50).numpy()
    ff, fb, mf, mb = orc.flow_stage(lf, lb, flow_up=flow_up, div_mode=div_mode)
    wf, wb = orc.sparse_corr(lf, lb, c1, c2, G, size, flow_up=flow_up, div_mode=div_mode)
    z = np.zeros((B, 1, G, G), np.float32)
    for wp, c, fl, mk in ((wf, c1, ff, mf), (wb, c2, fb, mb)):
        o = orc.regression_loss(z, z, c, c, 0.7, flow=None, size=size, want_grad=False, div_mode=div_mode)  # un-warped centres
        ox, oy, mg = orc.add_optical_flow(fl, o["cqx"].reshape(B, G, G), o["cqy"].reshape(B, G, G), size, mk, div_mode=div_mode)
        assert_bits_equal(wp[0], ox.reshape(B, -1), "warped x")
        assert_bits_equal(wp[1], oy.reshape(B, -1), "warped y")
        assert_bits_equal(wp[2] != 0, mg.reshape(B, -1), "mask bit")


CORR_TAGS = ["small", "raft_small", "oob"]


@pytest.mark.parametrize("tag", CORR_TAGS)
def test_corr_block(orc, tag):
    """the RAFT CorrBlock restatement against the reference's torch CorrBlock
    (contrast/flow/corr.py:12-60): volume within 1e-5, pyramid and windowed lookup bit-exact."""
    g = load_golden("corr_" + tag)
    L, r = int(g["num_levels"]), int(g["radius"])
    B, D, h, w = g["fmap1"].shape
    assert rel_err(orc.corr_volume(g["fmap1"], g["fmap2"]), g["level0"].reshape(B, h * w, h * w)) < TOL
    pyr = [g["level0"]]
    for l in range(1, L):
        pyr.append(orc.corr_pool(pyr[-1]))
        assert sha(pyr[-1]) == str(g[f"level{l}_sha"]), f"pyramid level {l}"
    assert_bits_equal(orc.corr_lookup(pyr, g["coords"], r), g["out"], "lookup")


@pytest.mark.parametrize("tag", ["small", "basic"])
def test_raft_state_dict_layout_matches_the_reference(tag):
    """Checkpoint compatibility of the drop-in estimator (contrast/flow/raft.py:26-162 + extractor.py + update.py): same
    state_dict keys, in the same order, with the same shapes as the reference's RAFT (recorded by the pin script)."""
    import types
    from contrast.flow import RAFT
    g = load_golden("raft_" + tag)
    m = RAFT(types.SimpleNamespace(small=bool(g["small"]), mixed_precision=False))
    sd = m.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    assert [",".join(map(str, v.shape)) for v in sd.values()] == [str(s) for s in g["shapes"]]
