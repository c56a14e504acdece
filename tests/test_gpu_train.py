"""Training slice: the backward kernels of the op-level drop-ins and the stage loss, called through
autograd over the C ABI, against the gradients of the live reference (train_ops.npz), against the CPU oracle on seeded
inputs, and at full size through the adjoint identity.

Tolerances: fp32 arithmetic; the scatter's float reductions arrive in a run-dependent order, so the bound is the fp32
re-association floor of a sum of <= a few hundred terms (a few 1e-6 relative to the largest gradient), plus the same
~1e-4 px coordinate noise as the forward test (fp64 coefficient inverse here, fp32 torch.inverse in the reference)."""
import pytest
import torch

import cds_mvsnet_b200 as C
from cds_mvsnet_b200 import losses, synthetic
from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _grad_enabled():
    """Other test modules switch autograd off process-wide at import; these tests need it."""
    with torch.enable_grad():
        yield


def cu(t):
    return t.to(DEV)


def close(a, b, atol, rtol=1e-5):
    torch.testing.assert_close(a.float().cpu(), b.float().cpu(), atol=atol, rtol=rtol)


# ------------------------------------------------------------------------------------------ A1 backward
def test_homo_warp_backward_golden(golden):
    g = golden("train_ops")
    for tag in ("planes", "pix"):
        fea = cu(g["warp_src_fea"]).requires_grad_(True)
        out = C.homo_warping_3D(fea, cu(g["warp_src_proj"]), cu(g["warp_ref_proj"]), cu(g[f"warp_depth_{tag}"]))
        assert out.requires_grad
        out.backward(cu(g["warp_grad_out"]))
        # gradient magnitudes reach ~30 (sums of up to ~50 unit-normal terms); coordinate noise moves a weight by ~1e-4
        close(fea.grad, g[f"warp_grad_src_{tag}"], 3e-3, 1e-4)
        assert O.rel_l1(fea.grad.cpu(), g[f"warp_grad_src_{tag}"]) < 2e-4


@pytest.mark.parametrize("Cc", [5, 8])   # 5: scalar reductions into NCHW; 8: vector reductions into the channels-last workspace
def test_homo_warp_backward_oracle_random(Cc):
    torch.manual_seed(3)
    s = synthetic.make_sample(dict(W=64, H=64, N=2, ndepths=(8,), ratios=(1.0,), B=3, Dtot=192, interval=2.65))
    pm = s.proj_matrices["stage1"]
    refP, srcP = O.compose_projection(pm[:, 0]), O.compose_projection(pm[:, 1])
    dv = 425 + 500 * torch.rand(3, 7, 16, 16)
    g_out = torch.randn(3, Cc, 7, 16, 16)
    fea = torch.randn(3, Cc, 16, 16, device=DEV, requires_grad=True)
    C.homo_warping_3D(fea, cu(srcP), cu(refP), cu(dv)).backward(cu(g_out))
    want = O.homo_warp_backward(g_out, srcP, refP, dv)
    assert O.rel_l1(fea.grad.cpu(), want) < 1e-4
    close(fea.grad, want, 1e-3, 1e-4)
    # no gradient reaches the cameras or the hypotheses (the reference builds the grid under no_grad, warping.py:79)
    fea2 = fea.detach().clone().requires_grad_(True)
    dvg, sp = cu(dv).requires_grad_(True), cu(srcP).requires_grad_(True)
    C.homo_warping_3D(fea2, sp, cu(refP), dvg).sum().backward()
    assert dvg.grad is None and sp.grad is None and fea2.grad is not None
    # inference calls stay on the plain path
    with torch.no_grad():
        assert not C.homo_warping_3D(fea, cu(srcP), cu(refP), cu(dv)).requires_grad


@pytest.mark.parametrize("shape", [(1, 16, 32, 256, 320), (1, 16, 32, 592, 800)])
def test_homo_warp_adjoint_full_size(shape):
    """<warp(x), g> = <x, warp^T(g)> at the stage-2 training shape and at cfg2's stage-2 shape (1600x1184 / 2, D = 32)."""
    B, Cc, D, h, w = shape
    torch.manual_seed(11)
    s = synthetic.make_sample(dict(W=2 * w, H=2 * h, N=2, ndepths=(48, 32, 8), ratios=(4.0, 1.5, 0.75), B=B, Dtot=192, interval=2.65))
    pm = s.proj_matrices["stage2"]
    refP, srcP = cu(O.compose_projection(pm[:, 0])), cu(O.compose_projection(pm[:, 1]))
    dv = 500 + 300 * torch.rand(B, D, h, w, device=DEV)
    x = torch.randn(B, Cc, h, w, device=DEV, requires_grad=True)
    g = torch.randn(B, Cc, D, h, w, device=DEV)
    out = C.homo_warping_3D(x, srcP, refP, dv)
    lhs = (out.detach().double() * g.double()).sum()
    out.backward(g)
    rhs = (x.detach().double() * x.grad.double()).sum()
    assert (out.detach() != 0).float().mean() > 0.1          # the views overlap: the identity is not trivially 0 = 0
    assert abs(lhs - rhs) <= 2e-6 * (out.detach().double() * g.double()).abs().sum()
    # linearity of the adjoint: warp^T(2g) = 2 warp^T(g) up to the reduction order
    x2 = x.detach().clone().requires_grad_(True)
    C.homo_warping_3D(x2, srcP, refP, dv).backward(2 * g)
    assert O.rel_l1(x2.grad, 2 * x.grad) < 1e-6


# ------------------------------------------------------------------------------------------ A5 backward
def test_depth_regression_backward_golden(golden):
    g = golden("train_ops")
    for tag in ("planes", "pix"):
        p = cu(g["regress_p"]).requires_grad_(True)
        dv = cu(g[f"warp_depth_{tag}"]).requires_grad_(True)
        C.depth_regression(p, dv).backward(cu(g["regress_grad_depth"]))
        close(p.grad, g[f"regress_grad_p_{tag}"], 0.0, 0.0)              # one product per element: exact
        close(dv.grad, g[f"regress_grad_dv_{tag}"], 2e-4 if tag == "planes" else 0.0, 1e-5)   # planes: a sum over h*w pixels
    # hypotheses without a gradient (the cascade's case, models/model.py:177-178 detaches them)
    p = cu(g["regress_p"]).requires_grad_(True)
    C.depth_regression(p, cu(g["warp_depth_pix"])).backward(cu(g["regress_grad_depth"]))
    close(p.grad, g["regress_grad_p_pix"], 0.0, 0.0)


def test_softargmin_chain_matches_torch_autograd():
    """softmax (torch) -> depth_regression (CUDA backward) against the same chain differentiated by torch alone."""
    torch.manual_seed(5)
    logits = torch.randn(2, 8, 24, 40, device=DEV)
    dv = 425 + 500 * torch.rand(2, 8, 24, 40, device=DEV)
    gd = torch.randn(2, 24, 40, device=DEV)
    a = logits.clone().requires_grad_(True)
    C.depth_regression(torch.softmax(a, 1), dv).backward(gd)
    b = logits.clone().requires_grad_(True)
    (torch.softmax(b, 1) * dv).sum(1).backward(gd)
    close(a.grad, b.grad, 1e-4, 1e-5)


# ------------------------------------------------------------------------------------------ loss
def loss_case(g, requires_grad=True):
    mk = (lambda t: cu(t).requires_grad_(True)) if requires_grad else cu
    inputs = {f"stage{i}": {"depth": mk(g[f"loss_in_stage{i}.depth"]), "norm_curv": mk(g[f"loss_in_stage{i}.norm_curv"])} for i in (1, 2, 3)}
    inputs["refined_depth"] = mk(g["loss_in_refined_depth"])
    gts = {f"stage{i}": cu(g[f"loss_gt_stage{i}"]) for i in (1, 2, 3, 4)}
    masks = {f"stage{i}": cu(g[f"loss_mask_stage{i}"]) for i in (1, 2, 3, 4)}
    return inputs, gts, masks


def test_final_loss_golden(golden):
    g = golden("train_ops")
    inputs, gts, masks = loss_case(g)
    total, dl = losses.final_loss(inputs, gts, masks, dlossw=g["loss_dlossw"].tolist(), depth_interval=cu(g["loss_interval"]))
    close(total, g["loss_total"], 1e-6, 2e-6)
    close(dl, g["loss_depth"], 1e-6, 2e-6)
    total.backward()
    for i in (1, 2, 3):
        close(inputs[f"stage{i}"]["depth"].grad, g[f"loss_grad_stage{i}.depth"], 1e-9, 1e-5)
        close(inputs[f"stage{i}"]["norm_curv"].grad, g[f"loss_grad_stage{i}.norm_curv"], 1e-9, 1e-5)
    close(inputs["refined_depth"].grad, g["loss_grad_refined_depth"], 1e-9, 1e-5)


def test_final_loss_edges(golden):
    g = golden("train_ops")
    inputs, gts, masks = loss_case(g, requires_grad=False)
    # no weights, no refined depth: the reference's `else` branch (losses.py:39-40)
    del inputs["refined_depth"]
    total, _ = losses.final_loss(inputs, gts, masks, depth_interval=cu(g["loss_interval"]))
    cpu_in = {k: {kk: vv.cpu() for kk, vv in v.items()} for k, v in inputs.items()}
    want, _ = O.final_loss(cpu_in, {k: v.cpu() for k, v in gts.items()}, {k: v.cpu() for k, v in masks.items()},
                           depth_interval=g["loss_interval"])
    close(total, want, 1e-6, 2e-6)
    # an empty mask gives NaN like the reference's mean over an empty selection
    masks["stage2"] = torch.zeros_like(masks["stage2"])
    total, _ = losses.final_loss(inputs, gts, masks, depth_interval=cu(g["loss_interval"]))
    assert torch.isnan(total)
    # mismatched feat_distance / mask shapes are refused
    inputs["stage1"]["feat_distance"] = torch.zeros(2, 5, 3, 3, device=DEV)
    inputs["stage1"]["feat_target"] = torch.zeros(2, 5, 3, 3, device=DEV)
    with pytest.raises(RuntimeError):
        losses.final_loss(inputs, gts, masks, depth_interval=cu(g["loss_interval"]))
    # CPU tensors are refused
    with pytest.raises(RuntimeError):
        losses.final_loss({k: {kk: vv.cpu() for kk, vv in v.items() if not kk.startswith("feat_")} for k, v in inputs.items()},
                          {k: v.cpu() for k, v in gts.items()}, {k: v.cpu() for k, v in masks.items()}, depth_interval=g["loss_interval"])


def test_final_loss_with_feat_term_golden(golden):
    """A training-mode output dict (feat_distance / feat_target present, models/model.py:94): value and gradients of the live
    reference's final_loss, including the 5x binary-cross-entropy term with its neg / pos weight."""
    g = golden("train_ops")
    inputs, gts, masks = loss_case(g)
    del inputs["refined_depth"]
    for i in (1, 2, 3):
        k = f"stage{i}"
        inputs[k]["feat_distance"] = cu(g[f"lossf_in_{k}.feat_distance"]).requires_grad_(True)
        inputs[k]["feat_target"] = cu(g[f"lossf_in_{k}.feat_target"])
        close(losses._FeatLossFn.apply(inputs[k]["feat_distance"].detach(), inputs[k]["feat_target"], masks[k]), g[f"lossf_value_{k}"], 2e-6, 2e-6)
    total, dl = losses.final_loss(inputs, gts, masks, dlossw=g["loss_dlossw"].tolist(), depth_interval=cu(g["loss_interval"]))
    close(total, g["lossf_total"], 1e-5, 2e-6)
    close(dl, g["lossf_depth"], 1e-6, 2e-6)
    total.backward()
    for i in (1, 2, 3):
        k = f"stage{i}"
        close(inputs[k]["feat_distance"].grad, g[f"lossf_grad_{k}.feat_distance"], 1e-8, 2e-5)
        close(inputs[k]["depth"].grad, g[f"lossf_grad_{k}.depth"], 1e-9, 1e-5)


def test_training_chain_vs_cpu_autograd():
    """One source view of the reference's training graph (models/model.py:44-49,85-91 + losses.py:14-23): warp -> similarity ->
    softmax -> soft-argmin -> masked smooth-L1.  On the GPU the warp, the regression and the loss are the CUDA kernels (forward and
    backward) and the glue is torch; on the CPU the whole graph is the oracle differentiated by torch.  Gradients w.r.t. both
    feature maps must agree."""
    torch.manual_seed(9)
    B, Cc, D, h, w = 2, 8, 6, 24, 32
    s = synthetic.make_sample(dict(W=4 * w, H=4 * h, N=2, ndepths=(8,), ratios=(1.0,), B=B, Dtot=192, interval=2.65))
    pm = s.proj_matrices["stage1"]
    refP, srcP = O.compose_projection(pm[:, 0]), O.compose_projection(pm[:, 1])
    dv = (torch.linspace(450, 900, D).reshape(1, D, 1, 1) + 20 * torch.rand(B, D, h, w)).contiguous()
    ref0, src0 = torch.tanh(torch.randn(B, Cc, h, w)), torch.tanh(torch.randn(B, Cc, h, w))
    gt = 450 + 450 * torch.rand(B, h, w)
    mask = (torch.rand(B, h, w) > 0.3).float()
    iv = torch.tensor([2.65, 2.5])

    def graph(ref, src, warp, regress, loss, mv):
        sim = (ref.unsqueeze(2) * warp(src, mv(srcP), mv(refP), mv(dv))).sum(1)
        depth = regress(torch.softmax(sim, 1), mv(dv))
        return loss(depth, mv(gt), mv(mask), mv(iv))

    rc, sc = ref0.clone().requires_grad_(True), src0.clone().requires_grad_(True)
    lc = graph(rc, sc, O.homo_warp, O.depth_regression, lambda d, g, m, i: O.stage_loss(d, g, m, i)[0], lambda t: t)
    lc.backward()
    rg, sg = cu(ref0).requires_grad_(True), cu(src0).requires_grad_(True)
    lg = graph(rg, sg, C.homo_warping_3D, C.depth_regression, lambda d, g, m, i: losses._StageLossFn.apply(d, None, g, m, i)[0], cu)
    lg.backward()
    close(lg, lc.detach(), 1e-5, 1e-5)
    assert O.rel_l1(rg.grad.cpu(), rc.grad) < 2e-4
    assert O.rel_l1(sg.grad.cpu(), sc.grad) < 2e-4


# ------------------------------------------------------------------------------------------ A4 in training mode
def _ref_costreg(in_ch, sd_prefix_sd):
    from oracle import ref_live
    import contextlib, io
    _, rmodule, _, _ = ref_live.load()
    with contextlib.redirect_stdout(io.StringIO()):
        ref = rmodule.CostRegNet(in_channels=in_ch, base_channels=8)
    ref.load_state_dict(sd_prefix_sd, strict=True)
    return ref


@pytest.mark.parametrize("stage,shape", [(2, (1, 8, 8, 16, 24)), (1, (2, 16, 16, 8, 16)), (0, (1, 32, 8, 24, 8))],
                         ids=["s3_c8", "s2_c16_b2", "s1_c32"])
def test_costregnet_training_matches_reference_autograd(pretrained_sd, stage, shape):
    """CostRegNet in training mode (BatchNorm3d batch statistics) against the live reference's module in training mode:
    logits, gradient of the input volume, gradients of all 31 parameters, and the updated running statistics."""
    from oracle import ref_live
    if not ref_live.available():
        pytest.skip("oracle/_ref/reference_models.zip not shipped (run build())")
    pre = f"cost_regularization.{stage}."
    sd = {k[len(pre):]: v.clone() for k, v in pretrained_sd.items() if k.startswith(pre)}
    ref = _ref_costreg(shape[1], sd).train()
    ours = C.CostRegNet(shape[1], 8).to(DEV)
    ours.load_state_dict(sd, strict=True)
    ours.train()
    torch.manual_seed(stage)
    x = torch.randn(*shape)
    gout = torch.randn(shape[0], 1, *shape[2:])
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    yr.backward(gout)
    xo = cu(x).requires_grad_(True)
    yo = ours(xo)
    assert yo.requires_grad and yo.shape == yr.shape
    yo.backward(cu(gout))
    scale = lambda t: t.abs().max().item() + 1e-12
    assert (yo.detach().cpu() - yr.detach()).abs().max() < 2e-4 * scale(yr)
    assert (xo.grad.cpu() - xr.grad).abs().max() < 5e-4 * scale(xr.grad)
    rp, op = dict(ref.named_parameters()), dict(ours.named_parameters())
    assert rp.keys() == op.keys() and len(rp) == 31
    for k in rp:
        assert op[k].grad is not None, k
        err = (op[k].grad.cpu() - rp[k].grad).abs().max().item()
        assert err < 1e-3 * scale(rp[k].grad) + 1e-6, (k, err, scale(rp[k].grad))
    rb, ob = dict(ref.named_buffers()), dict(ours.named_buffers())
    for k in rb:
        torch.testing.assert_close(ob[k].cpu().to(rb[k].dtype), rb[k], rtol=1e-4, atol=1e-5)
    # a second step: gradients accumulate into .grad like any autograd node's, running statistics keep moving
    ours(xo).backward(cu(gout))
    k = "conv6.conv.weight"
    assert (op[k].grad.cpu() - 2 * rp[k].grad).abs().max() < 2e-3 * scale(rp[k].grad)
    # eval mode afterwards runs the tensor-core inference path on the updated statistics
    ref(x.clone())                                  # the reference's second training step (running statistics)
    ours.eval(); ref.eval()
    with torch.no_grad():
        ye, yre = ours(cu(x)), ref(x)
    assert O.rel_l1(ye.cpu(), yre) < 5e-3           # fp16 tensor-core path against fp32


# ------------------------------------------------------------------------------------------ A6 in training mode
@pytest.mark.parametrize("layer,hw", [("conv00", (24, 40)), ("conv01", (19, 33)), ("conv10", (16, 24)), ("out1", (12, 20))])
def test_dynamic_conv_training_matches_reference_autograd(pretrained_sd, layer, hw):
    """DynamicConv in training mode against the live reference's module: outputs, gradient of the input, gradients of every
    parameter (branch convolutions, curvature convolutions, gate MLP, BatchNorm affine) and the gate's running statistics."""
    from oracle import ref_live
    from cds_mvsnet_b200 import weights as W
    if not ref_live.available():
        pytest.skip("oracle/_ref/reference_models.zip not shipped (run build())")
    _, _, rdyn, _ = ref_live.load()
    cin, cout, ks, pre = W.DYN_LAYERS[layer]
    sd = {k[len(pre) + 1:]: v.clone() for k, v in pretrained_sd.items() if k.startswith(pre + ".")}
    has_bias = any(k.endswith("convs.0.bias") for k in sd)
    ref = rdyn.DynamicConv(cin, cout, size_kernels=ks, bias=has_bias)
    ref.load_state_dict(sd, strict=True)
    ref.train()
    ours = C.DynamicConv(cin, cout, size_kernels=ks, bias=has_bias).to(DEV)
    ours.load_state_dict(sd, strict=True)
    ours.train()
    torch.manual_seed(len(layer) + hw[0])
    B = 2
    x = torch.randn(B, cin, *hw)
    epi = torch.tensor([[hw[1] * 1.3, -hw[0] * 0.4], [-25.0, hw[0] / 2.0]])
    gy, gn = torch.randn(B, cout, *hw), torch.randn(B, 1, *hw)
    T = 0.05
    xr = x.clone().requires_grad_(True)
    yr, nr = ref(xr, epi, T)
    (yr * gy).sum().add((nr * gn).sum()).backward()
    xo = cu(x).requires_grad_(True)
    yo, no = ours(xo, cu(epi), T)
    ((yo * cu(gy)).sum() + (no * cu(gn)).sum()).backward()
    scale = lambda t: t.abs().max().item() + 1e-12
    # the softmax over the branches at temperature T amplifies 1e-7 differences of the gate logits; outputs agree to ~1e-4
    assert (yo.detach().cpu() - yr.detach()).abs().max() < 5e-4 * scale(yr)
    assert (no.detach().cpu() - nr.detach()).abs().max() < 5e-4 * scale(nr)
    assert (xo.grad.cpu() - xr.grad).abs().max() < 2e-3 * scale(xr.grad)
    rp, op = dict(ref.named_parameters()), dict(ours.named_parameters())
    assert rp.keys() == op.keys()
    for k in rp:
        assert op[k].grad is not None, k
        err = (op[k].grad.cpu() - rp[k].grad).abs().max().item()
        assert err < 2e-3 * scale(rp[k].grad) + 1e-6, (k, err, scale(rp[k].grad))
    for k, v in ref.named_buffers():
        torch.testing.assert_close(dict(ours.named_buffers())[k].cpu().to(v.dtype), v, rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------------------------------ a whole training step
def test_training_step_through_patched_reference(pretrained_sd):
    """The reference's own CDSMVSNet in TRAINING mode (its forward with gt_depths, its StageNet with the feat_distance branch,
    its FeatureNet / Conv2d) with the leaf operators rebound by ``patch(level="leaf")`` -- DynamicConv and CostRegNet in their
    training forms, the differentiable warp and regression -- against the same model unpatched: outputs of every stage and the
    gradient of every one of its parameters after one backward."""
    from oracle import ref_live
    if not ref_live.available():
        pytest.skip("oracle/_ref/reference_models.zip not shipped (run build())")
    rmodel, rmodule, _, _ = ref_live.load()
    cfg = dict(W=96, H=64, N=3, ndepths=(8, 8, 8), ratios=(4.0, 2.0, 1.0), B=1, Dtot=48, interval=2.65 * 4)
    s = synthetic.make_sample(cfg, "plane", seed=3)
    imgs, dv = cu(s.imgs), cu(s.depth_values)
    proj = {k: cu(v) for k, v in s.proj_matrices.items()}
    gt_full = s.gt_depth if getattr(s, "gt_depth", None) is not None else None
    T = 0.05
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        def step(model):
            model.train()
            with torch.no_grad():
                model.eval()
                pseudo = model(imgs, proj, dv, temperature=T)
                gts = {f"stage{i}": pseudo[f"stage{i}"]["depth"].detach() * 1.01 for i in (1, 2, 3)}
                model.train()
            out = model(imgs, proj, dv, gt_depths=gts, temperature=T)
            torch.manual_seed(11)
            loss = 0.0
            for i in (1, 2, 3):
                o = out[f"stage{i}"]
                loss = loss + (o["depth"] * torch.rand_like(o["depth"])).mean() / 600.0
                loss = loss + (o["feat_distance"] * torch.rand_like(o["feat_distance"])).mean() + o["norm_curv"].mean()
            loss.backward()
            return out, loss.detach(), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}

        ref = ref_live.build_model(pretrained_sd, cfg["ndepths"], cfg["ratios"], device=DEV, rmodel=rmodel)
        out_r, loss_r, g_r = step(ref)
        saved = C.patch(rmodel, rmodule, level="leaf")
        try:
            ours = ref_live.build_model(pretrained_sd, cfg["ndepths"], cfg["ratios"], device=DEV, rmodel=rmodel)
            assert type(ours.cost_regularization[0]).__module__.startswith("cds_mvsnet_b200")
            out_o, loss_o, g_o = step(ours)
        finally:
            C.unpatch(saved)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    for i in (1, 2, 3):
        for name in ("depth", "feat_distance", "norm_curv"):
            a, b = out_o[f"stage{i}"][name].detach(), out_r[f"stage{i}"][name].detach()
            err = O.rel_l1(a.cpu(), b.cpu())
            print(f"train step stage{i}.{name}: rel-L1 {err:.2e}")
            assert err < 2e-3, (i, name, err)
    assert abs(float(loss_o) - float(loss_r)) < 1e-3 * abs(float(loss_r))
    assert g_o.keys() == g_r.keys() and len(g_r) > 200
    worst = max(((g_o[k] - g_r[k]).abs().max().item() / (g_r[k].abs().max().item() + 1e-9), k) for k in g_r)
    print("train step: worst parameter-gradient error (relative to the gradient's max):", worst)
    assert worst[0] < 5e-2, worst
    rel = torch.tensor([(g_o[k] - g_r[k]).norm().item() / (g_r[k].norm().item() + 1e-12) for k in g_r])
    assert rel.median() < 5e-3, rel.median()


def test_optimizer_steps_on_the_patched_reference(pretrained_sd):
    """A few Adam steps of the reference's training recipe on one sample -- its model (patched at "leaf"), this repository's
    fused final_loss, the temperature schedule -- reduce the loss; afterwards the SAME module objects in eval mode run the
    tensor-core inference kernels on the updated weights and running statistics (derived-weight caches follow the optimizer),
    and agree with the unpatched reference loaded with the trained state dict."""
    from oracle import ref_live
    if not ref_live.available():
        pytest.skip("oracle/_ref/reference_models.zip not shipped (run build())")
    rmodel, rmodule, _, _ = ref_live.load()
    cfg = dict(W=96, H=64, N=3, ndepths=(8, 8, 8), ratios=(4.0, 2.0, 1.0), B=1, Dtot=48, interval=2.65 * 4)
    s = synthetic.make_sample(cfg, "plane", seed=5)
    imgs, dv = cu(s.imgs), cu(s.depth_values)
    proj = {k: cu(v) for k, v in s.proj_matrices.items()}
    gt = cu(s.gt_depth)
    # the reference's forward always returns refined_depth (= the stage-3 depth without the refinement net), so its loss always
    # reads a "stage4" ground truth (models/losses.py:42-46)
    gts = {"stage1": gt[:, ::4, ::4].contiguous(), "stage2": gt[:, ::2, ::2].contiguous(), "stage3": gt, "stage4": gt}
    masks = {k: torch.ones_like(v) for k, v in gts.items()}
    saved = C.patch(rmodel, rmodule, level="leaf")
    try:
        model = ref_live.build_model(pretrained_sd, cfg["ndepths"], cfg["ratios"], device=DEV, rmodel=rmodel).train()
        opt = torch.optim.Adam(model.parameters(), lr=2e-4)
        history = []
        for it in range(4):
            opt.zero_grad()
            out = model(imgs, proj, dv, gt_depths=gts, temperature=losses.temperature_for_epoch(5))
            total, _ = losses.final_loss(out, gts, masks, dlossw=[0.5, 1.0, 2.0], depth_interval=cu(torch.tensor([cfg["interval"]])))
            assert torch.isfinite(total)
            total.backward()
            opt.step()
            history.append(float(total.detach()))
        print("training losses:", history)
        assert history[-1] < history[0]
        model.eval()
        with torch.no_grad():
            ours = model(imgs, proj, dv, temperature=0.01)
        trained = {k: v.detach().clone() for k, v in model.state_dict().items()}
    finally:
        C.unpatch(saved)
    ref = ref_live.build_model(trained, cfg["ndepths"], cfg["ratios"], device=DEV, rmodel=rmodel)
    with torch.no_grad():
        want = ref(imgs, proj, dv, temperature=0.01)
    for i in (1, 2, 3):
        err = O.rel_l1(ours[f"stage{i}"]["depth"].cpu(), want[f"stage{i}"]["depth"].cpu())
        assert err < 1e-3, (i, err)


# ------------------------------------------------------------------------------------------ the training kernels, op by op
@pytest.mark.parametrize("ci,co,shape,stride", [(8, 16, (5, 7, 9), 1), (3, 5, (6, 9, 130), 2), (20, 9, (4, 6, 7), 2), (1, 8, (3, 5, 300), 1)])
def test_train_conv3d_kernels_vs_torch(ci, co, shape, stride):
    """cds_train_conv3d / cds_train_conv3d_wgrad on odd shapes and channel counts (ragged channel tiles, partial strips, odd
    extents under stride 2) against torch in fp64."""
    from cds_mvsnet_b200 import train3d
    torch.manual_seed(ci + co)
    B = 2
    x, w = torch.randn(B, ci, *shape), torch.randn(co, ci, 3, 3, 3) * 0.2
    ref = torch.nn.functional.conv3d(x.double(), w.double(), stride=stride, padding=1)
    got = train3d.conv3d(cu(x), cu(train3d._tap_conv(w)), co, stride)
    assert got.shape == ref.shape
    assert (got.cpu().double() - ref).abs().max() < 1e-4 * ref.abs().max()
    g = torch.randn_like(ref).float()
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    torch.nn.functional.conv3d(xr, wr, stride=stride, padding=1).backward(g.double())
    dw = train3d.wgrad(cu(x), cu(g), stride).reshape(ci, 3, 3, 3, co).permute(4, 0, 1, 2, 3)
    assert (dw.cpu().double() - wr.grad).abs().max() < 2e-4 * wr.grad.abs().max()
    if stride == 1:   # input gradient = the same kernel with flipped, transposed weights
        dx = train3d.conv3d(cu(g), cu(w.flip(2, 3, 4).permute(0, 2, 3, 4, 1).reshape(co, 27, ci).contiguous()), ci, 1)
        assert (dx.cpu().double() - xr.grad).abs().max() < 1e-4 * xr.grad.abs().max()


@pytest.mark.parametrize("ci,co,shape", [(16, 8, (3, 4, 5)), (5, 12, (2, 3, 70))])
def test_train_deconv3d_kernels_vs_torch(ci, co, shape):
    """cds_train_deconv3d (ConvTranspose3d k3 s2 p1 op1), its input gradient (a stride-2 conv with the same weight) and its
    weight gradient (the wgrad kernel with input and gradient swapped)."""
    from cds_mvsnet_b200 import train3d
    torch.manual_seed(ci * co)
    B = 2
    x, w = torch.randn(B, ci, *shape), torch.randn(ci, co, 3, 3, 3) * 0.2
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    ref = torch.nn.functional.conv_transpose3d(xr, wr, stride=2, padding=1, output_padding=1)
    got = train3d.deconv3d(cu(x), cu(train3d._tap_deconv(w)), co)
    assert got.shape == ref.shape
    assert (got.cpu().double() - ref.detach()).abs().max() < 1e-4 * ref.abs().max()
    g = torch.randn_like(ref).float()
    ref.backward(g.double())
    dx = train3d.conv3d(cu(g), cu(w.permute(1, 2, 3, 4, 0).reshape(co, 27, ci).contiguous()), ci, 2)
    assert (dx.cpu().double() - xr.grad).abs().max() < 1e-4 * xr.grad.abs().max()
    dw = train3d.wgrad(cu(g), cu(x), 2).reshape(co, 3, 3, 3, ci).permute(4, 0, 1, 2, 3)
    assert (dw.cpu().double() - wr.grad).abs().max() < 2e-4 * wr.grad.abs().max()


@pytest.mark.parametrize("ci,co,k,hw", [(3, 11, 11, (20, 33)), (8, 19, 5, (9, 140)), (32, 35, 1, (7, 12)), (16, 4, 7, (11, 5))])
def test_train_conv2d_kernels_vs_torch(ci, co, k, hw):
    """Conv2dFn (cds_train_conv2d forward / input gradient, cds_train_conv2d_wgrad) for every kernel size of the feature
    extractor, images narrower than the kernel included."""
    from cds_mvsnet_b200.train2d import Conv2dFn
    torch.manual_seed(k + ci)
    x, w = torch.randn(2, ci, *hw), torch.randn(co, ci, k, k) / k
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    ref = torch.nn.functional.conv2d(xr, wr, padding=(k - 1) // 2)
    xo, wo = cu(x).requires_grad_(True), cu(w).requires_grad_(True)
    got = Conv2dFn.apply(xo, wo)
    assert (got.detach().cpu().double() - ref.detach()).abs().max() < 1e-4 * ref.abs().max()
    g = torch.randn_like(ref).float()
    ref.backward(g.double())
    got.backward(cu(g))
    assert (xo.grad.cpu().double() - xr.grad).abs().max() < 1e-4 * xr.grad.abs().max()
    assert (wo.grad.cpu().double() - wr.grad).abs().max() < 2e-4 * wr.grad.abs().max()
    with pytest.raises(RuntimeError):   # even kernel sizes are not part of the layer
        Conv2dFn.apply(xo, cu(torch.randn(co, ci, 4, 4)))
