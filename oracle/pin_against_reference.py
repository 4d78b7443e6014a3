"""Pin the CPU oracle against the REAL reference and write tests/golden/*.npz.

Run in the build container only (needs /root/reference; it cannot travel to the GPU box):

    python oracle/pin_against_reference.py            # check + (re)write goldens
    python oracle/pin_against_reference.py --check    # check only

What it does: imports the reference's own modules (contrast.util, contrast.flow,
contrast.models.PixPro) unmodified, with the three harness shims of SURVEY.md §8c
(gloo world_size=1; SyncBN->BN on CPU; Tensor.cuda no-op), runs them on seeded synthetic
inputs, asserts that oracle/pixpro_oracle.c reproduces every boolean/integer output
bit-exactly (FB masks, nearest-mask lookups, positive masks, pos_num), every coordinate
output bit-exactly (composite flows, up-sampled flows, warped grid centres) and every
float reduction within 1e-5 relative, then stores inputs + reference outputs as small
fixtures under tests/golden/.  Large dense outputs are stored as SHA-256 digests of their
raw bytes (bit-exactness makes a digest a sufficient golden).
"""
import argparse
import hashlib
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pixpro-with-opticalflow_b200"))
from oracle import oracle as orc  # noqa: E402
from pixpro_b200 import synth  # noqa: E402

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference():
    """Import the reference package under the name `contrast` without touching its sources."""
    for m in [m for m in sys.modules if m == "contrast" or m.startswith("contrast.")]:
        del sys.modules[m]
    sys.path.insert(0, REF)
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("gloo", rank=0, world_size=1)
    import contrast.util as rutil
    import contrast.flow as rflow
    import contrast.models  # noqa: F401
    rpix = sys.modules["contrast.models.PixPro"]  # the module (the package attribute is the class)
    import contrast.resnet as rresnet
    sys.path.remove(REF)
    torch.Tensor.cuda = lambda self, *a, **k: self  # util.py:196-197 hard-codes .cuda()
    return rutil, rflow, rpix, rresnet


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.dtype == np.float32:
        return a.shape == b.shape and np.array_equal(a.view(np.int32), b.view(np.int32))
    return a.shape == b.shape and np.array_equal(a, b)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


class Report:
    def __init__(self):
        self.rows = []
        self.ok = True

    def exact(self, name, got, want):
        ok = bits_equal(got, want)
        n_bad = int((np.asarray(got) != np.asarray(want)).sum()) if np.asarray(got).shape == np.asarray(want).shape else -1
        self.rows.append((name, "bit-exact" if ok else f"MISMATCH ({n_bad} elements)"))
        self.ok &= ok

    def close(self, name, got, want, tol=1e-5):
        e = rel_err(got, want)
        ok = e <= tol
        self.rows.append((name, f"rel {e:.2e} (tol {tol:g})" + ("" if ok else "  FAIL")))
        self.ok &= ok


def make_args(**kw):
    a = types.SimpleNamespace(alpha1=0.01, alpha2=0.5, use_flow_frames=False, use_flow_file=True, flow_up=True,
                              flow_cat_norm=False, debug=False, verbose=False)
    a.__dict__.update(kw)
    return a


def pixpro_args(**kw):
    a = types.SimpleNamespace(pixpro_p=2.0, pixpro_momentum=0.99, pixpro_pos_ratio=0.7, pixpro_clamp_value=0.0,
                              pixpro_transform_layer=1, pixpro_ins_loss_weight=0.0, output_dir="/tmp",
                              num_instances=1000, batch_size=4, epochs=10, start_epoch=1, feature_dim=256,
                              head_type="early_return")
    a.__dict__.update(kw)
    return a


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--only-corr", action="store_true", help="run only the RAFT CorrBlock section (fast) and write only its fixtures")
    opt = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    rutil, rflow, rpix, rresnet = import_reference()
    rep = Report()
    gold = {}
    if opt.only_corr:
        pin_corr_block(rep, gold)
        pin_raft(rep, gold, rflow)
        finish(rep, gold, opt)
        return

    # ---------------- a1 upflow8 ----------------
    fwd, bwd = synth.flow_fields(2, 3, h=18, w=32, seed=11)
    x = fwd.reshape(-1, 2, 18, 32)
    ref = rflow.upflow8(x).numpy()
    rep.exact("a1 upflow8 18x32", orc.upflow8(x.numpy()), ref)
    gold["upflow8"] = dict(inp=x.numpy(), out=ref)
    xf, _ = synth.flow_fields(1, 1, seed=12)
    ref = rflow.upflow8(xf.reshape(-1, 2, 90, 160)).numpy()
    rep.exact("a1 upflow8 90x160", orc.upflow8(xf.reshape(-1, 2, 90, 160).numpy()), ref)
    gold["upflow8_full"] = dict(seed=np.int64(12), inp=xf.numpy(), out_sha=np.array(sha(ref)))

    # ---------------- a2 normalise ----------------
    c = torch.randn(2, 2, 37, 53) * 40
    for name, rf, of in [("normalize_coord", rutil.normalize_coord, orc.normalize_coord),
                         ("normalize_flow", rutil.normalize_flow, orc.normalize_flow),
                         ("denormalize_flow", rutil.denormalize_flow, orc.denormalize_flow)]:
        ref = rf(c).numpy()
        rep.exact("a2 " + name, of(c.numpy()), ref)
        gold[name] = dict(inp=c.numpy(), out=ref)

    # ---------------- a3 concat_flow ----------------
    for tag, n, B, H, W, mag, is_norm in [("n1", 1, 2, 48, 64, 3.0, False), ("n2", 2, 2, 48, 64, 3.0, False),
                                          ("n5", 5, 2, 72, 128, 4.0, False), ("n5_oob", 5, 1, 40, 56, 25.0, False),
                                          ("n3_norm", 3, 2, 48, 64, 3.0, True), ("n1_norm", 1, 1, 24, 32, 3.0, True)]:
        f, _ = synth.flow_fields(B, n, h=H, w=W, magnitude=mag, seed=20 + n, coarse=(5, 7))
        flows = f.permute(1, 0, 2, 3, 4).contiguous()
        ref = rutil.concat_flow(flows, is_norm=is_norm).numpy()
        rep.exact(f"a3 concat_flow {tag}", orc.concat_flow(flows.numpy(), is_norm=is_norm), ref)
        gold[f"concat_flow_{tag}"] = dict(flows=flows.numpy(), is_norm=np.bool_(is_norm), out=ref)

    # ---------------- a5 forward_backward_consistency ----------------
    for tag, B, H, W, mag, is_norm in [("a", 2, 72, 128, 1.5, False), ("oob", 1, 40, 56, 30.0, False),
                                       ("norm", 1, 48, 64, 1.5, True)]:
        f, b = synth.flow_fields(B, 1, h=H, w=W, magnitude=mag, seed=31, coarse=(5, 7))
        f, b = f[:, 0].contiguous(), b[:, 0].contiguous()
        if is_norm:
            f, b = rutil.normalize_flow(f), rutil.normalize_flow(b)
        c0, c1, (m, cyc) = rutil.forward_backward_consistency(f, b, alpha_1=0.01, alpha_2=0.5, is_norm=is_norm)
        oc1, om, ocyc = orc.forward_backward_consistency(f.numpy(), b.numpy(), 0.01, 0.5, is_norm=is_norm)
        rep.exact(f"a5 fb mask {tag} (valid {m.float().mean():.3f})", om, m.numpy())
        rep.exact(f"a5 fb cycle {tag}", ocyc, cyc.numpy())
        rep.exact(f"a5 fb coords1 {tag}", oc1, c1.numpy())
        gold[f"fb_{tag}"] = dict(fwd=f.numpy(), bwd=b.numpy(), is_norm=np.bool_(is_norm), mask=m.numpy(),
                                 cycle=cyc.numpy(), coords1=c1.numpy())

    # ---------------- a6 apply_optical_flow (the flow stage) ----------------
    def run_apply(B, n, h, w, seed, args, mag=1.5):
        f, b = synth.flow_fields(B, n, h=h, w=w, seed=seed, magnitude=mag)
        H, W = (8 * h, 8 * w) if args.flow_up else (h, w)
        data = [None] * 7
        data[5] = [torch.zeros(B), f, b]
        data[6] = [torch.tensor([[H, W]] * B), torch.tensor([[n + 1]] * B)]
        (ff, size, mf), (fb_, _, mb) = rutil.apply_optical_flow(data, None, args)
        return f, b, ff, fb_, mf, mb

    for tag, B, n, h, w, kw in [("n1_up", 2, 1, 18, 32, {}), ("n5_up", 2, 5, 18, 32, {}),
                                ("n2_noup", 2, 2, 48, 64, dict(flow_up=False)),
                                ("n5_nomask", 1, 5, 18, 32, dict(alpha1=None, alpha2=None)),
                                ("n3_catnorm", 1, 3, 18, 32, dict(flow_cat_norm=True))]:
        args = make_args(**kw)
        f, b, ff, fb_, mf, mb = run_apply(B, n, h, w, 40 + n, args, mag=0.25 if args.flow_up else 1.5)
        off, ofb, omf, omb = orc.flow_stage(f.numpy(), b.numpy(), flow_up=args.flow_up, alpha_1=args.alpha1,
                                            alpha_2=args.alpha2, is_norm=args.flow_cat_norm)
        rep.exact(f"a6 flow_stage {tag} flow_fwd", off, ff.numpy())
        rep.exact(f"a6 flow_stage {tag} flow_bwd", ofb, fb_.numpy())
        g = dict(lo_fwd=f.numpy(), lo_bwd=b.numpy(), flow_up=np.bool_(args.flow_up),
                 use_mask=np.bool_(mf is not None), is_norm=np.bool_(args.flow_cat_norm),
                 flow_fwd=ff.numpy(), flow_bwd=fb_.numpy())
        if mf is not None:
            rep.exact(f"a6 flow_stage {tag} mask_fwd (valid {mf.float().mean():.3f})", omf, mf.numpy())
            rep.exact(f"a6 flow_stage {tag} mask_bwd", omb, mb.numpy())
            g.update(mask_fwd=np.packbits(mf.numpy()), mask_bwd=np.packbits(mb.numpy()))
            rep.close(f"a11 calc_mask_ratio {tag}", orc.calc_mask_ratio(mf.numpy()), rutil.calc_mask_ratio(mf).numpy(), 1e-6)
            g.update(mask_ratio_fwd=rutil.calc_mask_ratio(mf).numpy())
        gold[f"flow_stage_{tag}"] = g

    # full-size published shape (90x160 -> 720x1280), digests only
    for tag, n in [("full_n1", 1), ("full_n5", 5)]:
        args = make_args()
        f, b, ff, fb_, mf, mb = run_apply(1, n, 90, 160, 50 + n, args)
        off, ofb, omf, omb = orc.flow_stage(f.numpy(), b.numpy())
        rep.exact(f"a6 flow_stage {tag} flow_fwd", off, ff.numpy())
        rep.exact(f"a6 flow_stage {tag} flow_bwd", ofb, fb_.numpy())
        rep.exact(f"a6 flow_stage {tag} mask_fwd (valid {mf.float().mean():.3f})", omf, mf.numpy())
        rep.exact(f"a6 flow_stage {tag} mask_bwd", omb, mb.numpy())
        gold[f"flow_stage_{tag}"] = dict(seed=np.int64(50 + n), n=np.int64(n), lo_fwd=f.numpy(), lo_bwd=b.numpy(),
                                         flow_fwd_sha=np.array(sha(ff.numpy())), flow_bwd_sha=np.array(sha(fb_.numpy())),
                                         mask_fwd=np.packbits(mf.numpy()), mask_bwd=np.packbits(mb.numpy()))

    # ---------------- a7 add_optical_flow + a8 regression_loss ----------------
    def run_loss(tag, B, G, C, n, use_flow, use_mask, size_hw, flow_hw, seed, pos_ratio=0.7, mag=1.5):
        H, W = size_hw
        cq = synth.crop_coords(B, W, H, seed=seed)
        ck = synth.crop_coords(B, W, H, seed=seed + 1)
        gen = torch.Generator().manual_seed(seed)
        q = F.normalize(torch.randn(B, C, G, G, generator=gen), dim=1).requires_grad_(True)
        k = F.normalize(torch.randn(B, C, G, G, generator=gen), dim=1)
        flow = mask = None
        if use_flow:
            h, w = flow_hw[0] // 8, flow_hw[1] // 8
            f, b = synth.flow_fields(B, n, h=h, w=w, seed=seed + 2, magnitude=mag)
            off, ofb, omf, omb = orc.flow_stage(f.numpy(), b.numpy(), alpha_1=0.01 if use_mask else None,
                                                alpha_2=0.5 if use_mask else None)
            flow = torch.from_numpy(off)
            flow_b = torch.from_numpy(ofb)
            if use_mask:
                mask = torch.from_numpy(omf)
            size_t = torch.tensor([H, W])
            coord_q = [cq, [flow, size_t, mask]]
            coord_k = [ck, [flow_b, size_t, None]]
        else:
            coord_q, coord_k = cq, ck
        loss, (pos_num, pos_mean) = rpix.regression_loss(q, k, coord_q, coord_k, pos_ratio)
        loss.backward()
        o = orc.regression_loss(q.detach().numpy(), k.numpy(), cq.numpy(), ck.numpy(), pos_ratio,
                                flow=None if flow is None else flow.numpy(), size=(H, W),
                                mask=None if mask is None else mask.numpy())
        # the reference does not return pos_mask; recover it from its own centre arithmetic
        rep.exact(f"a8 {tag} pos_num {pos_num.tolist()[:4]}", o["pos_num"], pos_num.numpy())
        rep.close(f"a8 {tag} pos_mean", o["pos_mean"], pos_mean.numpy(), 1e-6)
        rep.close(f"a8 {tag} loss {loss.item():.6f}", o["loss"], loss.item())
        rep.close(f"a8 {tag} dq", o["dq"], q.grad.numpy())
        g = dict(q=q.detach().numpy(), k=k.numpy(), coord_q=cq.numpy(), coord_k=ck.numpy(),
                 pos_ratio=np.float64(pos_ratio), size=np.array([H, W]), loss=np.float32(loss.item()),
                 pos_num=pos_num.numpy(), pos_mean=pos_mean.numpy(), dq=q.grad.numpy(),
                 pos_mask=np.packbits(o["pos_mask"]),
                 near_threshold_pairs=orc.near_threshold_pairs(o, cq.numpy(), ck.numpy(), G, (H, W), pos_ratio))
        rep.rows.append((f"a8 {tag} pairs within 1e-5 of the threshold", f"{g['near_threshold_pairs'].tolist()} of {G ** 4} per sample"))
        if use_flow:
            # a7 directly
            P = G * G
            ox, oy, mg = rpix.add_optical_flow(flow, torch.from_numpy(_centres(o, cq, G, H, W)[0]).view(B, G, G),
                                               torch.from_numpy(_centres(o, cq, G, H, W)[1]).view(B, G, G),
                                               (H, W), mask)
            rep.exact(f"a7 {tag} out_x", o["cqx"].reshape(B, G, G), ox.numpy())
            rep.exact(f"a7 {tag} out_y", o["cqy"].reshape(B, G, G), oy.numpy())
            # the sparse restatement (flow stage evaluated at the grid centres only) against the reference's dense route
            swf, _ = orc.sparse_corr(f.numpy(), b.numpy(), cq.numpy(), None, G, (H, W), alpha_1=0.01 if use_mask else None,
                                     alpha_2=0.5 if use_mask else None)
            rep.exact(f"sparse {tag} out_x", swf[0].reshape(B, G, G), ox.numpy())
            rep.exact(f"sparse {tag} out_y", swf[1].reshape(B, G, G), oy.numpy())
            if mg is not None:
                rep.exact(f"sparse {tag} mask_grid", swf[2].reshape(mg.shape) != 0, mg.numpy())
            g.update(lo_fwd=f.numpy(), lo_bwd=b.numpy(), use_mask=np.bool_(use_mask), cqx=ox.numpy().reshape(B, P),
                     cqy=oy.numpy().reshape(B, P))
            if mg is not None:
                g.update(mask_grid=mg.numpy().reshape(B, P))
        gold[f"loss_{tag}"] = g

    def _centres(o, cq, G, H, W):
        # unwarped query centres, via the oracle's own centre routine on the no-flow path
        B = cq.shape[0]
        o2 = orc.regression_loss(np.zeros((B, 1, G, G), np.float32), np.zeros((B, 1, G, G), np.float32),
                                 cq.numpy(), cq.numpy(), 0.7, size=(H, W), want_grad=False)
        return o2["cqx"], o2["cqy"]

    run_loss("noflow_g7", 4, 7, 256, 0, False, False, (720, 1280), None, 60)
    run_loss("noflow_g14", 2, 14, 64, 0, False, False, (720, 1280), None, 61)
    run_loss("flow_g7_n1_mask", 4, 7, 256, 1, True, True, (720, 1280), (720, 1280), 62)
    run_loss("flow_g7_n5_mask", 2, 7, 64, 5, True, True, (720, 1280), (720, 1280), 63)
    run_loss("flow_g14_n2_nomask", 2, 14, 64, 2, True, False, (720, 1280), (720, 1280), 64)
    run_loss("flow_g7_diffsize", 2, 7, 32, 1, True, True, (720, 1280), (360, 640), 65)
    run_loss("flow_g7_big", 2, 7, 32, 1, True, True, (720, 1280), (720, 1280), 66, mag=12.0)
    run_loss("noflow_g7_ratio03", 3, 7, 32, 0, False, False, (720, 1280), None, 67, pos_ratio=0.3)
    # strip the big fixtures down: keep only hashes of inputs we can regenerate
    for key in list(gold):
        if key.startswith("loss_flow") and "lo_fwd" in gold[key]:
            pass

    # many random no-flow / flow mask cases -> pos_mask exactness via pos_num (cheap, no fixture)
    bad = 0
    tot = 0
    for s in range(200):
        cq = synth.crop_coords(8, seed=1000 + s)
        ck = synth.crop_coords(8, seed=5000 + s)
        q = torch.zeros(8, 1, 7, 7)
        loss, (pn, _) = rpix.regression_loss(q, q, cq, ck, 0.7)
        o = orc.regression_loss(q.numpy(), q.numpy(), cq.numpy(), ck.numpy(), 0.7, want_grad=False)
        bad += int((o["pos_num"] != pn.numpy()).sum())
        tot += 8
    rep.rows.append((f"a8 pos_num over {tot} random crop pairs", "bit-exact" if bad == 0 else f"MISMATCH in {bad}"))
    rep.ok &= bad == 0

    # ---------------- a9 featprop fwd/bwd through the reference module ----------------
    for tag, layer, p, cv, G, B in [("l1_p2_g7", 1, 2.0, 0.0, 7, 3), ("l0_p1_g7", 0, 1.0, 0.0, 7, 2),
                                    ("l1_p2_g14", 1, 2.0, 0.0, 14, 2), ("l0_p05_cv01", 0, 0.5, 0.1, 7, 2),
                                    ("l0_p3_g7", 0, 3.0, 0.0, 7, 2)]:
        m = rpix.PixPro.__new__(rpix.PixPro)
        torch.nn.Module.__init__(m)
        m.pixpro_p, m.pixpro_clamp_value = p, cv
        torch.manual_seed(70 + G)
        m.value_transform = rpix.conv1x1(256, 256) if layer == 1 else rpix.Identity()
        feat = torch.randn(B, 256, G, G, requires_grad=True)
        gout = torch.randn(B, 256, G, G)
        out = F.normalize(m.featprop(feat), dim=1)
        out.backward(gout)
        with torch.no_grad():
            val = m.value_transform(feat.detach())
        o = orc.featprop(feat.detach().numpy(), val.numpy(), gamma=p, clamp_value=cv)
        rep.close(f"a9 featprop {tag} fwd", o, out.detach().numpy())
        dfs, dv = orc.featprop_bwd(feat.detach().numpy(), val.numpy(), gout.numpy(), gamma=p, clamp_value=cv)
        dvt = torch.from_numpy(dv)
        if layer == 1:
            wt = m.value_transform.weight.detach()[:, :, 0, 0]
            dfeat = torch.from_numpy(dfs) + torch.einsum("oc,bohw->bchw", wt, dvt)
        else:
            dfeat = torch.from_numpy(dfs) + dvt
        rep.close(f"a9 featprop {tag} d_feat", dfeat.numpy(), feat.grad.numpy(), 2e-5)
        g = dict(feat=feat.detach().numpy(), val=val.numpy(), gout=gout.numpy(), gamma=np.float64(p),
                 clamp=np.float64(cv), out=out.detach().numpy(), d_feat=feat.grad.numpy())
        if layer == 1:
            g.update(weight=m.value_transform.weight.detach().numpy(), bias=m.value_transform.bias.detach().numpy(),
                     d_weight=m.value_transform.weight.grad.numpy(), d_bias=m.value_transform.bias.grad.numpy())
        gold[f"featprop_{tag}"] = g

    # ---------------- a4 / a6 general path: use_flow_frames (all sub-chains) and the debug return structure ----------------
    for tag, B, n, h, w, kw in [("frames_n3", 2, 3, 9, 16, dict(use_flow_frames=True)),
                                ("debug_n2", 2, 2, 9, 16, dict(debug=True))]:
        args = make_args(**kw)
        f, b = synth.flow_fields(B, n, h=h, w=w, seed=80 + n, magnitude=0.25)
        H, W = 8 * h, 8 * w
        data = [None] * 7
        data[5] = [torch.zeros(B), f, b]
        data[6] = [torch.tensor([[H, W]] * B), torch.tensor([[n + 1]] * B)]
        (ff, size, mf), (fb_, _, mb) = rutil.apply_optical_flow(data, None, args)
        g = dict(lo_fwd=f.numpy(), lo_bwd=b.numpy(), use_flow_frames=np.bool_(args.use_flow_frames), debug=np.bool_(args.debug),
                 flow_fwd=ff.numpy(), flow_bwd=fb_.numpy(), size=np.asarray(size))
        if args.debug:  # masks come back as [mask, cycle] lists (util.py:218-227)
            g.update(mask_fwd=np.packbits(mf[0].numpy()), mask_bwd=np.packbits(mb[0].numpy()), mask_shape=np.array(mf[0].shape),
                     cycle_fwd=mf[1].numpy(), cycle_bwd=mb[1].numpy())
            want = orc.flow_stage(f.numpy(), b.numpy())
            rep.exact(f"a6 general path {tag} flow_fwd", want[0], ff.numpy())
            rep.exact(f"a6 general path {tag} mask_fwd", want[2], mf[0].numpy())
        else:   # stacks over the n(n+1)/2 contiguous sub-chains (util.py:111-126), shortest first
            g.update(mask_fwd=np.packbits(mf.numpy()), mask_bwd=np.packbits(mb.numpy()), mask_shape=np.array(mf.shape))
            up_f = rflow.upflow8(f.permute(1, 0, 2, 3, 4).reshape(-1, 2, h, w)).reshape(n, B, 2, H, W).numpy()
            idx = 0
            for span in range(1, n + 1):
                for s0 in range(n - span + 1):
                    rep.exact(f"a4 all_concat_flow {tag} sub-chain [{s0}:{s0 + span}]", orc.concat_flow(up_f[s0:s0 + span]), ff[idx].numpy())
                    idx += 1
        gold[f"apply_general_{tag}"] = g

    pin_corr_block(rep, gold)
    pin_raft(rep, gold, rflow)

    # ---------------- a10 / a12 / cfg 0: the whole reference model, one fwd+bwd step on CPU ----------------
    # BASELINE configs[0]: PixPro ResNet-50, n_frames=1 (no flow), batch 4, 224x224 two-view crops, 7x7 grid.  Weights come
    # from synth.seeded_init_ (a function of each parameter's NAME, so the drop-in model can rebuild them on the GPU
    # box); inputs from seeded generators.  Stored: loss, pos_num, per-parameter gradient norms, a strided gradient sample.
    def ref_model_step(ins_weight, seed):
        torch.manual_seed(seed)
        m = rpix.PixPro(rresnet.resnet50, pixpro_args(pixpro_ins_loss_weight=ins_weight))
        for mod in list(m.modules()):   # SURVEY 8c shim 2: SyncBatchNorm has no CPU forward
            for cname, child in list(mod.named_children()):
                if isinstance(child, torch.nn.SyncBatchNorm):
                    bn = torch.nn.BatchNorm2d(child.num_features, child.eps, child.momentum, child.affine, child.track_running_stats)
                    bn.load_state_dict(child.state_dict())
                    setattr(mod, cname, bn)
        synth.seeded_init_(m, seed)
        m.train()
        gen = torch.Generator().manual_seed(seed)
        im1 = torch.randn(4, 3, 224, 224, generator=gen)
        im2 = torch.randn(4, 3, 224, 224, generator=gen)
        c1, c2 = synth.crop_coords(4, seed=seed + 1), synth.crop_coords(4, seed=seed + 2)
        loss, ((pn1, _), (pn2, _)) = m(im1, im2, c1, c2, is_update_momentum=False)
        loss.backward()
        names = [n_ for n_, p_ in m.named_parameters() if p_.grad is not None]
        gn = np.array([float(dict(m.named_parameters())[n_].grad.double().norm()) for n_ in names])
        sample = np.concatenate([dict(m.named_parameters())[n_].grad.flatten()[::997][:8].numpy() for n_ in names[:40]])
        return dict(seed=np.int64(seed), ins_weight=np.float64(ins_weight), loss=np.float64(loss.item()), pos_num_1=pn1.numpy(),
                    pos_num_2=pn2.numpy(), grad_names=np.array(names), grad_norms=gn, grad_sample=sample)

    for tag, insw in [("cfg0", 0.0), ("cfg0_ins", 1.0)]:
        g = ref_model_step(insw, 90)
        rep.rows.append((f"a10/a12 reference model step {tag}", f"loss {float(g['loss']):.6f}, pos_num {g['pos_num_1'].tolist()}, "
                         f"{len(g['grad_names'])} gradient tensors"))
        gold[f"model_{tag}"] = g

    # ---------------- §8(f) rank 1: EMA of the key branch, LARS + SGD ----------------
    # (the reference's own contrast/lars.py, loaded by path: the package import would pull in termcolor)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_lars", os.path.join(REF, "contrast", "lars.py"))
    rlars = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rlars)
    torch.manual_seed(123)
    k0, q0 = torch.randn(3001), torch.randn(3001)
    mom = 1. - (1. - 0.99) * (np.cos(np.pi * 7 / 100) + 1) / 2.   # PixPro.py:326 at k=7, K=100
    ref = (k0 * mom + q0 * (1. - mom)).numpy()                    # PixPro.py:330
    rep.exact("f1 EMA update", orc.ema_update(k0.numpy(), q0.numpy(), mom), ref)
    gold["ema"] = dict(k=k0.numpy(), q=q0.numpy(), m=np.float64(mom), out=ref)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.ReLU(), torch.nn.Conv2d(8, 4, 1))
    groups = rlars.add_weight_decay(net, 1e-5)
    ropt = rlars.LARS(torch.optim.SGD(groups, lr=0.3, momentum=0.9), eps=1e-8, trust_coef=0.001)
    plist = [p for g_ in ropt.param_groups for p in g_["params"]]
    meta = [(g_["weight_decay"], g_["lr"], g_["momentum"], g_["dampening"], not g_["ignore"]) for g_ in ropt.param_groups
            for _ in g_["params"]]
    state = [(p.detach().numpy().copy(), None) for p in plist]
    lg = dict(n_params=np.int64(len(plist)), n_steps=np.int64(3), trust=np.float64(0.001), eps=np.float64(1e-8),
              meta=np.array([[m_[0], m_[1], m_[2], m_[3], float(m_[4])] for m_ in meta], np.float64))
    for i, (p0, _) in enumerate(state):
        lg[f"p{i}_init"] = p0
    for step in range(3):
        ropt.zero_grad()
        net(torch.randn(4, 3, 8, 8)).square().mean().backward()
        grads = [p.grad.detach().numpy().copy() for p in plist]
        ropt.step()
        for i, p in enumerate(plist):
            wd, lr, mo, da, lars = meta[i]
            pn, bn, _ = orc.lars_sgd_step(state[i][0], grads[i], state[i][1], wd, lr, mo, da, lars=lars, first=state[i][1] is None)
            want = p.detach().numpy().copy()
            if lars:
                rep.close(f"f1 LARS step {step} tensor {i}", pn, want, 1e-6)
            else:
                rep.exact(f"f1 SGD  step {step} tensor {i} (LARS-ignored)", pn, want)
            state[i] = (pn, bn)
            lg[f"g{i}_s{step}"] = grads[i]
            lg[f"p{i}_s{step}"] = want
    gold["lars_sgd"] = lg

    finish(rep, gold, opt)


def pin_corr_block(rep, gold):
    """the reference's torch CorrBlock (contrast/flow/corr.py, loaded by path: the package import is
    already done, this only avoids `alt_cuda_corr`'s try/except noise) against orc_corr_volume / _pool / _lookup."""
    import importlib
    rcorr = importlib.import_module("contrast.flow.corr")
    for tag, B, D, h, w, L, r, spread in [("small", 2, 32, 8, 12, 3, 2, 3.0), ("raft_small", 1, 128, 16, 24, 4, 3, 6.0),
                                          ("oob", 1, 16, 8, 8, 2, 4, 20.0)]:
        g = torch.Generator().manual_seed(700 + h)
        f1 = torch.randn(B, D, h, w, generator=g)
        f2 = torch.randn(B, D, h, w, generator=g)
        from contrast.flow.utils.utils import coords_grid
        coords = coords_grid(B, h, w) + spread * torch.randn(B, 2, h, w, generator=g)
        blk = rcorr.CorrBlock(f1, f2, num_levels=L, radius=r)
        ref_out = blk(coords).numpy()
        ref_pyr = [p.numpy() for p in blk.corr_pyramid]
        vol = orc.corr_volume(f1.numpy(), f2.numpy())
        rep.close(f"f4 corr volume {tag}", vol, ref_pyr[0].reshape(B, h * w, h * w), 1e-5)
        # pyramid / lookup from the REFERENCE's level 0 (bit-exact stages are then checked bit-exactly)
        pyr = [ref_pyr[0]]
        for l in range(1, L):
            pyr.append(orc.corr_pool(pyr[-1]))
            rep.exact(f"f4 corr pyramid {tag} level {l}", pyr[-1], ref_pyr[l])
        out = orc.corr_lookup(pyr, coords.numpy(), r)
        rep.exact(f"f4 corr lookup {tag} (L={L}, r={r})", out, ref_out)
        gold[f"corr_{tag}"] = dict(fmap1=f1.numpy(), fmap2=f2.numpy(), coords=coords.numpy(), num_levels=np.int64(L), radius=np.int64(r),
                                   level0=ref_pyr[0], out=ref_out, **{f"level{l}_sha": np.array(sha(ref_pyr[l])) for l in range(1, L)})


def pin_raft(rep, gold, rflow):
    """SURVEY 8(f) rank 4, the caller side: the reference's RAFT estimator (contrast/flow/raft.py:26-162, small and basic
    variants) run on CPU with name-seeded weights (synth.seeded_init_) in eval mode; the drop-in contrast.flow.RAFT rebuilds
    the same weights on the GPU box and must reproduce the flows (tests/test_gpu_corr.py).  The stored key list pins the
    state_dict layout (checkpoint compatibility) on the CPU side."""
    for tag, small, B, H, W, iters in [("small", True, 2, 128, 160, 4), ("basic", False, 1, 128, 192, 2)]:
        args = argparse.Namespace(small=small, mixed_precision=False)
        m = rflow.RAFT(args)
        synth.seeded_init_(m, 300)
        m.eval()
        g = torch.Generator().manual_seed(310 + int(small))
        base = torch.rand(B, 3, H, W, generator=g) * 255.0
        # stored as float16: rounded BEFORE the reference sees them, so that the fixture holds exactly what it was given
        im1 = base.half().float()
        im2 = (torch.roll(base, shifts=(2, -3), dims=(2, 3)) + 4.0 * torch.rand(B, 3, H, W, generator=g)).half().float()
        with torch.no_grad():
            low, up = m(im1, im2, iters=iters, upsample=False, test_mode=True)
        sd = m.state_dict()
        rep.rows.append((f"f4 reference RAFT ({tag}, {iters} iterations, {H}x{W})",
                         f"{len(sd)} state_dict entries, |flow| max {float(low.abs().max()):.3f} low-res px"))
        gold[f"raft_{tag}"] = dict(small=np.bool_(small), iters=np.int64(iters), seed=np.int64(300), image1=im1.numpy().astype(np.float16),
                                   image2=im2.numpy().astype(np.float16), flow_low=low.numpy(), flow_up_sha=np.array(sha(up.numpy())),
                                   flow_up_absmax=np.float64(up.abs().max()), keys=np.array(list(sd.keys())),
                                   shapes=np.array([",".join(map(str, v.shape)) for v in sd.values()]))


def finish(rep, gold, opt):
    width = max(len(r[0]) for r in rep.rows)
    for name, res in rep.rows:
        print(f"{name:<{width}}  {res}")
    print("ORACLE PIN:", "OK" if rep.ok else "FAILED")
    if not rep.ok:
        sys.exit(1)
    if not opt.check:
        os.makedirs(GOLD, exist_ok=True)
        for key, g in gold.items():
            np.savez_compressed(os.path.join(GOLD, key + ".npz"), **g)
        tot = sum(os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD))
        print(f"wrote {len(gold)} fixtures, {tot / 1e6:.2f} MB, to {GOLD}")


if __name__ == "__main__":
    main()
