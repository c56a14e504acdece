"""Parity tests proper: every CUDA op, called through the C ABI / drop-in surface, against the golden
vectors generated from the live reference and against the CPU oracle on the same seeded inputs.

Tolerances: the fp32-storage build of each kernel must sit at the fp32 re-association floor (the
reference itself moves by ~1e-6 relative between thread counts, SURVEY.md 8c fixture 5); the fp16-
storage production setting is bounded by fp16 rounding of operands (2^-11 relative per element) and
checked end-to-end against north_star's 1e-3 relative-L1 bar in test_gpu_e2e.py.
"""
import pytest
import torch

import cds_mvsnet_b200 as C
from cds_mvsnet_b200 import _lib, synthetic, weights as W
from cds_mvsnet_b200._lib import call, ptr
from oracle import oracle as O

pytestmark = pytest.mark.gpu
T = 0.01
DEV = "cuda"
torch.set_grad_enabled(False)


def cu(t):
    return t.to(DEV)


def close(a, b, atol, rtol=1e-5):
    torch.testing.assert_close(a.float().cpu(), b.float().cpu(), atol=atol, rtol=rtol)


def to_blocked(x):
    """NCDHW -> the library's channel-blocked 3-D layout [B, C/8, D, H, W, 8]."""
    B, C, D, H, Wd = x.shape
    return x.reshape(B, C // 8, 8, D, H, Wd).permute(0, 1, 3, 4, 5, 2).contiguous()


def from_blocked(x):
    B, C8, D, H, Wd, _ = x.shape
    return x.permute(0, 1, 5, 2, 3, 4).reshape(B, C8 * 8, D, H, Wd)


# ---------------------------------------------------------------------------------------- A1
def test_homo_warp_golden(golden):
    g = golden("warp")
    out = C.homo_warping_3D(cu(g["src_fea"]), cu(g["src_proj"]), cu(g["ref_proj"]), cu(g["depth_planes"]))
    # coefficients come from an fp64 inverse here, an fp32 torch.inverse there: ~1e-4 px of coordinate noise
    close(out, g["out_planes"], 2e-4)
    out = C.homo_warping_3D(cu(g["src_fea"]), cu(g["src_proj"]), cu(g["ref_proj"]), cu(g["depth_pix"]))
    close(out, g["out_pix"], 2e-4)
    assert out[:, :, 0].abs().max() == 0          # far-out-of-frustum plane: zero padding


def test_homo_warp_oracle_random():
    torch.manual_seed(1)
    from cds_mvsnet_b200 import synthetic
    s = synthetic.make_sample(dict(W=64, H=64, N=2, ndepths=(8,), ratios=(1.0,), B=3, Dtot=192, interval=2.65))
    pm = s.proj_matrices["stage1"]
    fea = torch.randn(3, 5, 16, 16)          # odd channel count: the op-level kernel has no C % 8 restriction
    dv = 425 + 500 * torch.rand(3, 7, 16, 16)
    refP, srcP = O.compose_projection(pm[:, 0]), O.compose_projection(pm[:, 1])
    close(C.homo_warping_3D(cu(fea), cu(srcP), cu(refP), cu(dv)), O.homo_warp(fea, srcP, refP, dv), 2e-5)


def test_homo_warp_rejects_bad_shapes():
    with pytest.raises(RuntimeError):
        C.homo_warping_3D(torch.zeros(1, 8, 4, 4, device=DEV), torch.eye(4, device=DEV)[None], torch.eye(4, device=DEV)[None],
                          torch.ones(1, 2, 5, 5, device=DEV))


# ---------------------------------------------------------------------------------------- A8
def test_camera_setup_golden(golden):
    g = golden("epipole")
    B = g["cam_ref"].shape[0]
    pm = cu(torch.stack((g["cam_ref"], g["cam_src"]), 1).contiguous())   # [B,2,2,4,4]
    coef = torch.empty(1, B, 1, 12, device=DEV)
    epi = torch.empty(2, 1, B, 2, device=DEV)
    import ctypes
    call("cds_camera_setup", (ctypes.c_void_p * 1)(pm.data_ptr()), 1, 0, B, 2, ptr(coef), ptr(epi))
    close(epi[0, 0], g["e_ref"], 2e-2, 2e-5)
    close(epi[1, 0], g["e_src"], 2e-2, 2e-5)
    rot, trans = O.warp_coefficients(O.compose_projection(g["cam_src"]), O.compose_projection(g["cam_ref"]))
    close(coef[0, :, 0, :9].reshape(B, 3, 3), rot, 1e-5, 1e-5)
    close(coef[0, :, 0, 9:], trans, 1e-3, 1e-5)


# ---------------------------------------------------------------------------------------- A9
@pytest.mark.parametrize("tag,D,ratio,scale", [("s1", 48, 4.0, 4), ("s2", 32, 1.5, 2), ("s3", 8, 0.75, 1)])
def test_hypotheses_golden(golden, tag, D, ratio, scale):
    g = golden("hypotheses")
    dv = cu(g["depth_values"])
    ref = g[f"{tag}_out"]
    B, _, h, w = ref.shape
    out = torch.empty(B, D, h, w, device=DEV)
    prev = None if tag == "s1" else cu(g[f"{tag}_prev"]).contiguous()
    hp, wp = (0, 0) if prev is None else prev.shape[1:]
    call("cds_depth_hypotheses", ptr(dv), dv.shape[1], ptr(prev), hp, wp, B, D, ratio, h * scale, w * scale, scale, ptr(out))
    close(out, ref, 5e-4, 0)


# ---------------------------------------------------------------------------------------- A5
def test_tail_golden(golden):
    g = golden("tail")
    logits, dsm = cu(g["logits"]), cu(g["depth_samples"])
    B, D, h, w = logits.shape
    depth, conf, prob = (torch.empty(B, h, w, device=DEV), torch.empty(B, h, w, device=DEV), torch.empty(B, D, h, w, device=DEV))
    call("cds_softmax_regress", ptr(logits), ptr(dsm), 1, 0, B, D, h, w, ptr(depth), ptr(conf), ptr(prob))
    close(prob, torch.softmax(g["logits"], 1), 1e-6)
    close(depth, g["depth"], 2e-3, 1e-6)
    # the window index is a truncation: allow the rare pixel whose expectation sits on an integer
    bad = ((conf.cpu() - g["conf"]).abs() > 1e-5).float().mean()
    assert bad < 0.002
    # drop-in functions on probabilities
    p = torch.softmax(logits, 1)
    close(C.depth_regression(p, dsm), g["depth"], 2e-3, 1e-6)
    assert ((C.conf_regression(p).cpu() - g["conf"]).abs() > 1e-5).float().mean() < 0.002
    planes = torch.linspace(400, 900, D, device=DEV)[None].repeat(B, 1)
    close(C.depth_regression(p, planes), O.depth_regression(p.cpu(), planes.cpu()), 2e-3, 1e-6)


# ---------------------------------------------------------------------------------------- A6
@pytest.mark.parametrize("storage,atol", [(torch.float32, 2e-4), (torch.float16, 2e-2)])
@pytest.mark.parametrize("name", ["conv00", "conv01", "conv10", "conv20", "out1", "out3"])
def test_dynamic_conv_golden(golden, pretrained_sd, name, storage, atol):
    g = golden("dynconv")
    cin, cout, ks, pre = W.DYN_LAYERS[name]
    m = C.DynamicConv(cin, cout, size_kernels=ks, bias=name.startswith("out"), storage=storage)
    m.load_state_dict({k[len(pre) + 1:]: v for k, v in pretrained_sd.items() if k.startswith(pre + ".")})
    m = m.to(DEV).eval()
    y, nc = m(cu(g[f"{name}_x"]), epipole=cu(g[f"{name}_epi"]), temperature=T)
    ref_y, ref_nc = g[f"{name}_y"], g[f"{name}_nc"]
    # the gate is softmax(g / 0.01): compare in relative-L1 (a handful of pixels sit on a gate edge)
    assert O.rel_l1(y.cpu(), ref_y) < (2e-5 if storage == torch.float32 else 3e-3)
    assert O.rel_l1(nc.cpu(), ref_nc) < (2e-5 if storage == torch.float32 else 3e-3)
    assert (y.cpu() - ref_y).abs().max() < 50 * atol


def test_dynamic_conv_refuses_stride_and_switches_to_its_training_form():
    with pytest.raises(NotImplementedError):
        C.DynamicConv(8, 8, stride=2)
    m = C.DynamicConv(8, 8).to(DEV)          # a fresh module is in training mode: the autograd form (tests/test_gpu_train.py)
    with torch.enable_grad():
        y, nc = m(torch.zeros(1, 8, 16, 16, device=DEV), epipole=torch.zeros(1, 2, device=DEV))
    assert y.requires_grad and y.shape == (1, 8, 16, 16) and nc.shape == (1, 1, 16, 16)
    with pytest.raises(TypeError):
        m(torch.zeros(1, 8, 16, 16, device=DEV))                       # the epipole is not optional
    # the fused modules above it stay inference-only
    f = C.FeatureNet(8).to(DEV)
    with pytest.raises(NotImplementedError):
        f(torch.zeros(1, 3, 32, 32, device=DEV), epipole=torch.zeros(1, 2, device=DEV))


# ---------------------------------------------------------------------------------------- A7
@pytest.mark.parametrize("storage,tol", [(torch.float32, 5e-5), (torch.float16, 4e-3)])
def test_feature_net_golden(golden, pretrained_sd, storage, tol):
    g = golden("featurenet")
    m = C.FeatureNet(8, storage=storage)
    m.load_state_dict({k[8:]: v for k, v in pretrained_sd.items() if k.startswith("feature.")})
    m = m.to(DEV).eval()
    out = m(cu(g["img"]), epipole=cu(g["epi"]), temperature=T)
    for st in ("stage1", "stage2", "stage3"):
        fea, ncs, nca = out[st]
        assert fea.shape == g[f"{st}_fea"].shape and ncs.shape == g[f"{st}_nc_sum"].shape
        assert O.rel_l1(fea.cpu(), g[f"{st}_fea"]) < tol, st
        assert O.rel_l1(ncs.cpu(), g[f"{st}_nc_sum"]) < 4 * tol, st
        assert O.rel_l1(nca.cpu(), g[f"{st}_nc_abs"]) < 2 * tol, st


def test_feature_net_batch_equals_single(pretrained_sd):
    """Several images with different epipoles in one launch == one at a time (the cascade batches 2(N-1) images)."""
    torch.manual_seed(3)
    m = C.FeatureNet(8, storage=torch.float32)
    m.load_state_dict({k[8:]: v for k, v in pretrained_sd.items() if k.startswith("feature.")})
    m = m.to(DEV).eval()
    img = torch.rand(3, 3, 64, 96, device=DEV)
    epi = torch.tensor([[250.0, -40.0], [-500.0, 30.0], [48.0, 32.0]], device=DEV)
    both = m(img, epipole=epi, temperature=T)
    keep = {k: tuple(t.clone() for t in v) for k, v in both.items()}
    for i in range(3):
        one = m(img[i:i + 1], epipole=epi[i:i + 1], temperature=T)
        for st in keep:
            for a, b in zip(keep[st], one[st]):
                close(a[i:i + 1], b, 1e-5, 1e-5)


# ---------------------------------------------------------------------------------------- A3 / A4
@pytest.mark.parametrize("st", [0, 1, 2])
def test_visnet_golden(golden, pretrained_sd, st):
    g = golden("nets3d")
    x = cu(g[f"vis{st}_x"])
    n, _, h, w = x.shape
    wp = W.pack_visnet(pretrained_sd, f"stage_net.vis.{st}", DEV)
    out = torch.empty(n, h, w, device=DEV)
    ent, cur = x[:, 0].contiguous(), x[:, 1].contiguous()   # keep the temporaries alive across the launch
    call("cds_visnet", ptr(ent), ptr(cur), ptr(wp), n, h, w, ptr(out))
    close(out.unsqueeze(1), g[f"vis{st}_y"], 5e-6)


def test_visnet_tile_borders(pretrained_sd):
    """Sizes that are not multiples of the 32x16 tile, and several maps per launch."""
    torch.manual_seed(5)
    x = torch.cat((2.5 * torch.rand(3, 1, 37, 75), 0.3 * torch.rand(3, 1, 37, 75)), 1)
    wp = W.pack_visnet(pretrained_sd, "stage_net.vis.1", DEV)
    out = torch.empty(3, 37, 75, device=DEV)
    xc = cu(x)
    ent, cur = xc[:, 0].contiguous(), xc[:, 1].contiguous()
    call("cds_visnet", ptr(ent), ptr(cur), ptr(wp), 3, 37, 75, ptr(out))
    close(out.unsqueeze(1), O.vis_net(x, pretrained_sd, "stage_net.vis.1"), 5e-6)


@pytest.mark.parametrize("storage,tol", [(torch.float32, 2e-5), (torch.float16, 5e-3)])
@pytest.mark.parametrize("st", [0, 1, 2])
def test_costregnet_golden(golden, pretrained_sd, st, storage, tol):
    g = golden("nets3d")
    cin = (32, 16, 8)[st]
    m = C.CostRegNet(cin, 8, storage=storage)
    pre = f"cost_regularization.{st}."
    m.load_state_dict({k[len(pre):]: v for k, v in pretrained_sd.items() if k.startswith(pre)})
    m = m.to(DEV).eval()
    y = m(cu(g[f"cr{st}_x"]))
    ref = g[f"cr{st}_y"]
    assert y.shape == ref.shape
    assert O.rel_l1(y.cpu(), ref) < tol


def test_costregnet_rejects_indivisible():
    m = C.CostRegNet(8, 8).to(DEV).eval()
    with pytest.raises(RuntimeError, match="divisible by 8"):
        m(torch.zeros(1, 8, 8, 12, 16, device=DEV))


@pytest.mark.parametrize("cin,cout,stride", [(8, 16, 2), (16, 16, 1), (64, 64, 1), (32, 64, 2)])
def test_conv3d_block_vs_torch(cin, cout, stride):
    """Single Conv3d block against the published operator (odd sizes exercise ceil(n/2) and borders)."""
    torch.manual_seed(cin + cout)
    x = torch.randn(2, cin, 5, 9, 11)
    w = torch.randn(cout, cin, 3, 3, 3) / (27 * cin) ** 0.5
    b = torch.randn(cout)
    ref = torch.relu(torch.nn.functional.conv3d(x, w, b, stride=stride, padding=1))
    xc = cu(to_blocked(x))
    wc = cu(w.permute(2, 3, 4, 1, 0).reshape(27, cin, cout).contiguous())
    out = torch.empty_like(cu(to_blocked(ref)))
    bc = cu(b)
    call("cds_conv3d_k3", ptr(xc), ptr(wc), ptr(bc), 2, cin, cout, 5, 9, 11, stride, 1, _lib.CDS_F32, ptr(out))
    close(from_blocked(out), ref, 2e-5, 1e-4)


@pytest.mark.parametrize("cin,cout", [(16, 8), (32, 16), (64, 32)])
def test_deconv3d_block_vs_torch(cin, cout):
    torch.manual_seed(cin)
    x = torch.randn(2, cin, 3, 5, 7)
    w = torch.randn(cin, cout, 3, 3, 3) / (8 * cin) ** 0.5
    b = torch.randn(cout)
    skip = torch.randn(2, cout, 6, 10, 14)
    ref = skip + torch.relu(torch.nn.functional.conv_transpose3d(x, w, b, stride=2, padding=1, output_padding=1))
    xc = cu(to_blocked(x))
    wc = cu(w.permute(2, 3, 4, 0, 1).reshape(27, cin, cout).contiguous())
    sc = cu(to_blocked(skip))
    out = torch.empty_like(sc)
    bc = cu(b)
    call("cds_deconv3d_k3s2", ptr(xc), ptr(wc), ptr(bc), ptr(sc), 2, cin, cout, 3, 5, 7, _lib.CDS_F32, ptr(out))
    close(from_blocked(out), ref, 2e-5, 1e-4)


# ---------------------------------------------------------------------------------------- A2
@pytest.mark.parametrize("C_", [8, 16, 32])
@pytest.mark.parametrize("storage,tol", [(torch.float32, 1e-5), (torch.float16, 2e-3)])
@pytest.mark.parametrize("V", [3, 6])   # <= 4 views: reference chunks held in registers; more: re-read per plane
def test_costvol_vs_oracle(C_, storage, tol, V):
    from cds_mvsnet_b200 import synthetic
    torch.manual_seed(C_)
    B, D, h, w = 2, 6, 24, 40
    s = synthetic.make_sample(dict(W=4 * w, H=4 * h, N=V + 1, ndepths=(8,), ratios=(1.0,), B=B, Dtot=192, interval=2.65))
    pm = s.proj_matrices["stage1"]
    ref_fea = torch.tanh(torch.randn(V, B, C_, h, w))
    src_fea = torch.tanh(torch.randn(V, B, C_, h, w))
    if storage == torch.float16:       # compare like with like: the oracle sees the fp16-rounded features
        ref_fea, src_fea = ref_fea.half().float(), src_fea.half().float()
    dv = 425 + 500 * torch.rand(B, D, h, w)
    vis = torch.rand(V, B, h, w)
    refP = O.compose_projection(pm[:, 0])
    ents, vol = [], 0.0
    for v in range(V):
        warped = O.homo_warp(src_fea[v], O.compose_projection(pm[:, v + 1]), refP, dv)
        prod, ent = O.similarity_entropy(ref_fea[v], warped)
        ents.append(ent[:, 0])
        vol = vol + prod * vis[v].unsqueeze(1).unsqueeze(1)
    vol = vol / (vis.sum(0).unsqueeze(1).unsqueeze(1) + 1e-6)
    import ctypes
    pmc = cu(pm)
    coef = torch.empty(1, B, V, 12, device=DEV)
    call("cds_camera_setup", (ctypes.c_void_p * 1)(pmc.data_ptr()), 1, 0, B, V + 1, ptr(coef), None)
    rf = cu(ref_fea.permute(0, 1, 3, 4, 2).contiguous()).to(storage)
    sf = cu(src_fea.permute(0, 1, 3, 4, 2).contiguous()).to(storage)
    dvc, visc = cu(dv), cu(vis)
    ent_out = torch.empty(V, B, h, w, device=DEV)
    dt = _lib.dtype_code(storage)
    call("cds_costvol_entropy", ptr(rf), ptr(sf), ptr(coef), ptr(dvc), V, B, C_, D, h, w, dt, ptr(ent_out))
    close(ent_out, torch.stack(ents), 2e-4, 1e-4)
    # the visibility net's copy: packed-half blend on fp16 features (entropy to ~5e-4), the same kernel otherwise
    ent_fast = torch.full_like(ent_out, float("nan"))
    call("cds_costvol_entropy_fast", ptr(rf), ptr(sf), ptr(coef), ptr(dvc), V, B, C_, D, h, w, dt, ptr(ent_fast))
    if storage == torch.float16:
        close(ent_fast, torch.stack(ents), 1e-3, 5e-4)
        assert (ent_fast - ent_out).abs().max() < 2e-3
    else:
        assert torch.equal(ent_fast, ent_out)
    vol_out = torch.empty(B, C_ // 8, D, h, w, 8, device=DEV, dtype=storage)
    call("cds_costvol_aggregate", ptr(rf), ptr(sf), ptr(coef), ptr(dvc), ptr(visc), V, B, C_, D, h, w, dt, ptr(vol_out))
    assert O.rel_l1(from_blocked(vol_out.float().cpu()), vol) < tol


@pytest.mark.parametrize("V", [2, 4])
@pytest.mark.parametrize("hw", [(27, 45), (120, 200)], ids=lambda v: f"{v[0]}x{v[1]}")
@pytest.mark.parametrize("regime", ["smooth", "noisy", "step"])
def test_costvol_stage3_tma_staged_vs_oracle(V, hw, regime):
    """The last stage's shape (C = 8, D = 8, fp16) runs the TMA-staged form of both sweeps: source boxes in shared memory when
    a tile's samples stay together ("smooth": hypotheses around a smooth depth map), the global gathers when they do not
    ("noisy": independent depths per pixel), and a mix of both inside one launch ("step": a depth discontinuity)."""
    from cds_mvsnet_b200 import synthetic
    import ctypes
    torch.manual_seed(V + hw[0])
    B, D, C_ = 2, 8, 8
    h, w = hw
    s = synthetic.make_sample(dict(W=4 * w, H=4 * h, N=V + 1, ndepths=(8,), ratios=(1.0,), B=B, Dtot=192, interval=2.65))
    pm = s.proj_matrices["stage1"]
    ref_fea = torch.tanh(torch.randn(V, B, C_, h, w)).half().float()
    src_fea = torch.tanh(torch.randn(V, B, C_, h, w)).half().float()
    ys, xs = torch.meshgrid(torch.arange(h).float(), torch.arange(w).float(), indexing="ij")
    centre = 600 + 0.8 * xs - 0.5 * ys
    if regime == "noisy":
        centre = 450 + 450 * torch.rand(h, w)
    elif regime == "step":
        centre = torch.where(xs > w / 2, centre + 150.0, centre)
    dv = (centre.reshape(1, 1, h, w) + 2.0 * (torch.arange(D).float() - 3.5).reshape(1, D, 1, 1)).repeat(B, 1, 1, 1).contiguous()
    vis = torch.rand(V, B, h, w)
    refP = O.compose_projection(pm[:, 0])
    ents, vol = [], 0.0
    for v in range(V):
        warped = O.homo_warp(src_fea[v], O.compose_projection(pm[:, v + 1]), refP, dv)
        prod, ent = O.similarity_entropy(ref_fea[v], warped)
        ents.append(ent[:, 0])
        vol = vol + prod * vis[v].unsqueeze(1).unsqueeze(1)
    vol = vol / (vis.sum(0).unsqueeze(1).unsqueeze(1) + 1e-6)
    pmc = cu(pm)
    coef = torch.empty(1, B, V, 12, device=DEV)
    call("cds_camera_setup", (ctypes.c_void_p * 1)(pmc.data_ptr()), 1, 0, B, V + 1, ptr(coef), None)
    rf = cu(ref_fea.permute(0, 1, 3, 4, 2).contiguous()).half()
    sf = cu(src_fea.permute(0, 1, 3, 4, 2).contiguous()).half()
    dvc, visc = cu(dv), cu(vis)
    ent_out = torch.full((V, B, h, w), float("nan"), device=DEV)
    call("cds_costvol_entropy", ptr(rf), ptr(sf), ptr(coef), ptr(dvc), V, B, C_, D, h, w, _lib.CDS_F16, ptr(ent_out))
    close(ent_out, torch.stack(ents), 2e-4, 1e-4)
    vol_out = torch.full((B, 1, D, h, w, 8), float("nan"), device=DEV, dtype=torch.float16)
    call("cds_costvol_aggregate", ptr(rf), ptr(sf), ptr(coef), ptr(dvc), ptr(visc), V, B, C_, D, h, w, _lib.CDS_F16, ptr(vol_out))
    torch.cuda.synchronize()
    assert not torch.isnan(vol_out.float()).any()
    assert O.rel_l1(from_blocked(vol_out.float().cpu()), vol) < 2e-3


@pytest.mark.parametrize("shape", [(8, 24, 40), (48, 37, 150), (32, 64, 300)], ids=lambda v: "x".join(map(str, v)))
def test_prob_head_fused_tail_matches_separate(pretrained_sd, shape):
    """cds_prob_head_regress (softmax over D + depth / confidence regression in the prob head's epilogue) against the rolling
    prob head followed by cds_softmax_regress: same arithmetic (logits and depth bit-identical); and both against torch."""
    D, H, Wd = shape
    torch.manual_seed(D)
    B = 2
    x = torch.relu(torch.randn(B, 8, D, H, Wd))
    wp = pretrained_sd["cost_regularization.2.prob.weight"]
    l = W.Conv3dWeights(8, 1, wp.permute(2, 3, 4, 1, 0).reshape(27, 8, 1).contiguous().to(DEV), torch.zeros(1, device=DEV))
    packed = W.pack_conv3d_roll(l)
    xb = cu(to_blocked(x)).half()
    samples = cu(450 + 30 * torch.arange(D).float().reshape(1, D, 1, 1) + 5 * torch.rand(B, D, H, Wd)).contiguous()
    logits_a = torch.empty(B, D, H, Wd, device=DEV); logits_b = torch.empty_like(logits_a)
    depth_a = torch.empty(B, H, Wd, device=DEV); conf_a = torch.empty_like(depth_a)
    depth_b = torch.full((B, H, Wd), float("nan"), device=DEV); conf_b = torch.full_like(depth_b, float("nan"))
    call("cds_conv3d_k3_roll", ptr(xb), ptr(packed), None, B, 8, 1, D, H, Wd, 0, ptr(logits_a))
    call("cds_softmax_regress", ptr(logits_a), ptr(samples), 1, 0, B, D, H, Wd, ptr(depth_a), ptr(conf_a), None)
    call("cds_prob_head_regress", ptr(xb), ptr(packed), ptr(samples), B, D, H, Wd, ptr(logits_b), ptr(depth_b), ptr(conf_b))
    torch.cuda.synchronize()
    assert torch.equal(logits_a, logits_b)
    assert torch.equal(depth_a, depth_b)
    torch.testing.assert_close(conf_a, conf_b, rtol=1e-6, atol=1e-7)   # the window sum may contract into FMAs differently
    ref_l = torch.nn.functional.conv3d(xb.float().cpu().permute(0, 1, 5, 2, 3, 4).reshape(B, 8, D, H, Wd), wp, padding=1)[:, 0]
    pr = torch.softmax(ref_l, 1)
    assert O.rel_l1(depth_b.cpu(), O.depth_regression(pr, samples.cpu())) < 1e-4
    assert (conf_b.cpu() - O.conf_regression(pr)).abs().mean() < 2e-3


def test_costvol_rejects_too_many_views():
    z = torch.zeros(16, device=DEV)
    with pytest.raises(RuntimeError, match="V="):
        call("cds_costvol_entropy", ptr(z), ptr(z), ptr(z), ptr(z), 9, 1, 8, 4, 8, 8, 1, ptr(z))


# ---------------------------------------------------------------------------------------- A4 on tensor cores
@pytest.mark.parametrize("cin,cout", [(8, 8), (16, 8), (32, 8), (16, 16), (32, 32)])
@pytest.mark.parametrize("shape", [(3, 5, 133), (2, 9, 256), (1, 4, 128), (4, 6, 50), (2, 3, 100)])
def test_conv3d_tcgen05_vs_torch(cin, cout, shape):
    """tcgen05 implicit-GEMM Conv3d block against the published operator on fp16-rounded operands
    (fp32 accumulation on both sides, so only summation order differs)."""
    D, H, Wd = shape
    torch.manual_seed(cin * 100 + cout + Wd)
    x = torch.randn(2, cin, D, H, Wd).half().float()
    w = (torch.randn(cout, cin, 3, 3, 3) / (27 * cin) ** 0.5).half().float()
    b = torch.randn(cout)
    ref = torch.relu(torch.nn.functional.conv3d(x, w, b, stride=1, padding=1))
    lw = W.Conv3dWeights(cin, cout, w.permute(2, 3, 4, 1, 0).reshape(27, cin, cout).contiguous(), b)
    packed = cu(W.pack_conv3d_tc(lw))
    lib = _lib.LIB.load()
    assert lib.cds_conv3d_k3_tc_supported(cin, cout, D, H, Wd, 1) == 1
    assert packed.numel() == lib.cds_conv3d_k3_tc_weight_halfs(cin, cout)
    xc = cu(to_blocked(x)).half()
    bc = cu(b)
    out = torch.empty(2, cout // 8, D, H, Wd, 8, device=DEV, dtype=torch.float16)
    call("cds_conv3d_k3_tc", ptr(xc), ptr(packed), ptr(bc), 2, cin, cout, D, H, Wd, 1, ptr(out))
    torch.cuda.synchronize()
    got = from_blocked(out.float().cpu())
    # output is rounded to fp16: half an ulp at |y| <= 8 is 4e-3
    close(got, ref, 6e-3, 2e-3)


def test_conv3d_tcgen05_unsupported_shapes():
    lib = _lib.LIB.load()
    assert lib.cds_conv3d_k3_tc_supported(8, 8, 8, 8, 4, 1) == 0        # too narrow
    assert lib.cds_conv3d_k3_tc_supported(8, 16, 8, 8, 256, 2) == 0     # stride 2: CUDA-core kernel
    z = torch.zeros(64, device=DEV, dtype=torch.float16)
    with pytest.raises(RuntimeError, match="unsupported"):
        call("cds_conv3d_k3_tc", ptr(z), ptr(z), ptr(z), 1, 8, 24, 4, 4, 64, 1, ptr(z))


# ---------------------------------------------------------------------------------------- A6 on tensor cores
@pytest.mark.parametrize("name", ["conv00", "conv01", "out3", "conv10", "out2", "conv20", "out1"])
@pytest.mark.parametrize("hw", [(37, 200), (64, 128), (16, 333), (24, 40), (30, 100)])
def test_dynamic_conv_tcgen05_vs_oracle(pretrained_sd, name, hw):
    """tcgen05 DynamicConv (every layer shape of the feature extractor) against the oracle and the CUDA-core kernel."""
    cin, cout, ks, pre = W.DYN_LAYERS[name]
    torch.manual_seed(hash(name) % 1000 + hw[1])
    x = (torch.rand(2, cin, *hw) if cin == 3 else torch.randn(2, cin, *hw)).half().float()
    epi = torch.tensor([[hw[1] * 1.7, -hw[0] * 0.6], [-30.0, hw[0] / 2.0]])
    sd = {k[len(pre) + 1:]: v for k, v in pretrained_sd.items() if k.startswith(pre + ".")}
    ref_y, ref_nc = O.dynamic_conv(x, pretrained_sd, pre, ks, epi, T)
    outs = {}
    for use_tc in (True, False):
        m = C.DynamicConv(cin, cout, size_kernels=ks, bias=name.startswith("out"), storage=torch.float16, use_tc=use_tc)
        m.load_state_dict(sd)
        m = m.to(DEV).eval()
        y, nc = m(cu(x), epipole=cu(epi), temperature=T)
        torch.cuda.synchronize()
        outs[use_tc] = (y.cpu(), nc.cpu())
        assert O.rel_l1(y.cpu(), ref_y) < 4e-3, (use_tc, O.rel_l1(y.cpu(), ref_y))
        assert O.rel_l1(nc.cpu(), ref_nc) < 4e-3, (use_tc, O.rel_l1(nc.cpu(), ref_nc))
    assert O.rel_l1(outs[True][0], outs[False][0]) < 4e-3


@pytest.mark.parametrize("V,B,hw", [(4, 1, (24, 150)), (2, 2, (17, 130)), (5, 1, (8, 40)), (1, 1, (9, 128))])
def test_dynamic_conv_pairs_matches_plain(pretrained_sd, V, B, hw):
    """conv00 over the cascade's (side, v, b) pair batch with the reference image's convolutions shared between its V pairs
    (cds_dynamic_conv_tc_pairs) against the plain per-item entry on the same batch: same MMAs, agreement to fp32 re-association."""
    import ctypes
    cin, cout, ks, pre = W.DYN_LAYERS["conv00"]
    H, Wd = hw
    torch.manual_seed(V * 10 + B)
    N = V + 1
    imgs = torch.rand(B * N, 3, H, Wd)
    n = 2 * V * B
    idx = torch.empty(2, V, B, dtype=torch.int32)
    for v in range(V):
        for b in range(B):
            idx[0, v, b], idx[1, v, b] = b * N, b * N + v + 1
    epi = torch.randn(n, 2) * Wd
    w = W.pack_dynamic_conv(pretrained_sd, pre, cin, cout, ks, DEV)
    w.tc = W.pack_dynamic_conv_tc(w)
    img8 = torch.empty(B * N, H, Wd, 8, device=DEV, dtype=torch.float16)
    imgs_c, idx_c, epi_c = cu(imgs), cu(idx.reshape(-1)), cu(epi)
    call("cds_image_to_nhwc8", ptr(imgs_c), B * N, H, Wd, ptr(img8))
    kz = (ctypes.c_int * 3)(*ks)
    res = []
    for pairs in (False, True):
        out = torch.full((n, H, Wd, cout), float("nan"), device=DEV, dtype=torch.float16)
        stats = torch.zeros(n, cout, 2, device=DEV, dtype=torch.float64)
        nc = torch.full((n, H, Wd), float("nan"), device=DEV)
        ncsq = torch.full((n, H, Wd), float("nan"), device=DEV)
        if pairs:
            call("cds_dynamic_conv_tc_pairs", ptr(img8), B * N, ptr(idx_c), ptr(epi_c), 1.0, ptr(w.tc), ptr(w.bias), ptr(w.gate), V, B,
                 8, cout, H, Wd, 3, kz, T, ptr(out), None, ptr(stats), ptr(nc), ptr(ncsq), 0, None)
        else:
            call("cds_dynamic_conv_tc", ptr(img8), B * N, ptr(idx_c), None, 0, ptr(epi_c), 1.0, ptr(w.tc), ptr(w.bias), ptr(w.gate), n,
                 8, cout, H, Wd, 3, kz, T, 0, ptr(out), None, ptr(stats), ptr(nc), ptr(ncsq), 0, None)
        torch.cuda.synchronize()
        res.append((out.float().cpu(), stats.cpu(), nc.cpu(), ncsq.cpu()))
    # same MMAs; the two epilogues add the weight-residual product and the bias in a different association, so the fp32
    # results agree to rounding and the stored fp16 values to one ulp
    torch.testing.assert_close(res[0][0], res[1][0], rtol=2e-3, atol=1e-5)
    assert (res[0][0] != res[1][0]).float().mean() < 0.02
    torch.testing.assert_close(res[0][2], res[1][2], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(res[0][3], res[1][3], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(res[0][1], res[1][1], rtol=1e-6, atol=1e-3)
    # and both follow the oracle
    x_items = torch.stack([imgs[int(i)] for i in idx.reshape(-1)])
    ref_y, _ = O.dynamic_conv(x_items, pretrained_sd, pre, ks, epi, T)
    assert O.rel_l1(res[1][0].permute(0, 3, 1, 2), ref_y) < 4e-3


def test_prob_head_tcgen05_vs_torch():
    """8 -> 1 prob conv on the tensor cores (hi + residual fp16 weights) against the published operator."""
    torch.manual_seed(11)
    D, H, Wd = 8, 10, 200
    x = torch.randn(2, 8, D, H, Wd).half().float()
    w = torch.randn(1, 8, 3, 3, 3) / 14.7
    ref = torch.nn.functional.conv3d(x, w, padding=1)[:, 0]
    lw = W.Conv3dWeights(8, 1, w.permute(2, 3, 4, 1, 0).reshape(27, 8, 1).contiguous(), torch.zeros(1))
    packed = cu(W.pack_conv3d_tc(lw))
    xc = cu(to_blocked(x)).half()
    out = torch.empty(2, D, H, Wd, device=DEV)
    call("cds_conv3d_k3_tc", ptr(xc), ptr(packed), None, 2, 8, 1, D, H, Wd, 0, ptr(out))
    close(out, ref, 2e-4, 1e-4)


@pytest.mark.parametrize("cin,cout", [(16, 8), (32, 16)])
@pytest.mark.parametrize("shape", [(3, 5, 133), (2, 4, 256), (1, 2, 128), (3, 4, 25), (2, 2, 100)])
def test_deconv3d_tcgen05_vs_torch(cin, cout, shape):
    """tcgen05 transposed conv (8 output parity classes side by side in N) against the published operator."""
    D, H, Wd = shape
    torch.manual_seed(cin + Wd)
    x = torch.randn(2, cin, D, H, Wd).half().float()
    w = (torch.randn(cin, cout, 3, 3, 3) / (8 * cin) ** 0.5).half().float()
    b = torch.randn(cout)
    skip = torch.randn(2, cout, 2 * D, 2 * H, 2 * Wd).half().float()
    ref = skip + torch.relu(torch.nn.functional.conv_transpose3d(x, w, b, stride=2, padding=1, output_padding=1))
    lw = W.Conv3dWeights(cin, cout, w.permute(2, 3, 4, 0, 1).reshape(27, cin, cout).contiguous(), b)
    packed = cu(W.pack_deconv3d_tc(lw))
    lib = _lib.LIB.load()
    assert lib.cds_deconv3d_k3s2_tc_supported(cin, cout, D, H, Wd) == 1
    assert packed.numel() == lib.cds_deconv3d_k3s2_tc_weight_halfs(cin, cout)
    xc, sc, bc = cu(to_blocked(x)).half(), cu(to_blocked(skip)).half(), cu(b)
    out = torch.empty_like(sc)
    call("cds_deconv3d_k3s2_tc", ptr(xc), ptr(packed), ptr(bc), ptr(sc), 2, cin, cout, D, H, Wd, ptr(out))
    torch.cuda.synchronize()
    close(from_blocked(out.float().cpu()), ref, 8e-3, 2e-3)


@pytest.mark.parametrize("cin,cout,stride", [(8, 16, 2), (16, 16, 1), (16, 32, 2), (32, 32, 1), (32, 64, 2), (64, 64, 1),
                                             (8, 16, 1), (32, 32, 2)])
@pytest.mark.parametrize("shape", [(5, 9, 11), (8, 16, 24), (2, 3, 133), (6, 37, 50)])
def test_conv3d_gather_tcgen05_vs_torch(cin, cout, stride, shape):
    """Gather-form tcgen05 Conv3d block (stride 1|2, weights fed as fp16 value + residual columns) against the published
    operator on fp16-rounded activations and UNROUNDED weights; odd sizes exercise ceil(n/2), borders and partial tiles."""
    D, H, Wd = shape
    torch.manual_seed(cin * 100 + cout + Wd + stride)
    x = torch.randn(2, cin, D, H, Wd).half().float()
    w = torch.randn(cout, cin, 3, 3, 3) / (27 * cin) ** 0.5
    b = torch.randn(cout)
    ref = torch.relu(torch.nn.functional.conv3d(x, w, b, stride=stride, padding=1))
    lw = W.Conv3dWeights(cin, cout, w.permute(2, 3, 4, 1, 0).reshape(27, cin, cout).contiguous(), b)
    packed = cu(W.pack_conv3d_gtc(lw))
    lib = _lib.LIB.load()
    assert lib.cds_conv3d_k3_gtc_supported(cin, cout, stride) == 1
    assert packed.numel() == lib.cds_conv3d_k3_gtc_weight_halfs(cin, cout)
    xc, bc = cu(to_blocked(x)).half(), cu(b)
    out = torch.full(tuple(to_blocked(ref).shape), float("nan"), device=DEV, dtype=torch.float16)
    call("cds_conv3d_k3_gtc", ptr(xc), ptr(packed), ptr(bc), 2, cin, cout, D, H, Wd, stride, 1, ptr(out))
    torch.cuda.synchronize()
    # output is rounded to fp16: half an ulp at |y| <= 8 is 4e-3
    close(from_blocked(out.float().cpu()), ref, 6e-3, 2e-3)


@pytest.mark.parametrize("cin,cout", [(8, 8), (16, 8), (32, 8), (8, 1)])
@pytest.mark.parametrize("shape", [(3, 5, 133), (8, 16, 256), (1, 4, 128), (4, 6, 50), (5, 19, 300), (2, 3, 126), (9, 40, 127)])
def test_conv3d_rolling_tcgen05_vs_torch(cin, cout, shape):
    """Persistent d-rolling tcgen05 Conv3d (kw folded into N, weights as fp16 value + residual columns) against the
    published operator on fp16-rounded activations and UNROUNDED weights; shapes cover partial / overlapping x tiles,
    H not a multiple of the row tile, single planes and several columns per CTA."""
    D, H, Wd = shape
    torch.manual_seed(cin * 100 + cout + Wd)
    x = torch.randn(2, cin, D, H, Wd).half().float()
    w = torch.randn(cout, cin, 3, 3, 3) / (27 * cin) ** 0.5
    b = torch.randn(cout) if cout > 1 else torch.zeros(1)
    ref = torch.nn.functional.conv3d(x, w, b if cout > 1 else None, stride=1, padding=1)
    lw = W.Conv3dWeights(cin, cout, w.permute(2, 3, 4, 1, 0).reshape(27, cin, cout).contiguous(), b)
    packed = cu(W.pack_conv3d_roll(lw))
    lib = _lib.LIB.load()
    assert lib.cds_conv3d_k3_roll_supported(cin, cout, D, H, Wd) == 1
    assert packed.numel() == lib.cds_conv3d_k3_roll_weight_halfs(cin, cout)
    xc, bc = cu(to_blocked(x)).half(), cu(b)
    if cout == 1:
        out = torch.full((2, D, H, Wd), float("nan"), device=DEV, dtype=torch.float32)
        call("cds_conv3d_k3_roll", ptr(xc), ptr(packed), None, 2, cin, 1, D, H, Wd, 0, ptr(out))
        torch.cuda.synchronize()
        close(out.cpu(), ref[:, 0], 2e-4, 1e-4)
    else:
        ref = torch.relu(ref)
        out = torch.full((2, 1, D, H, Wd, 8), float("nan"), device=DEV, dtype=torch.float16)
        call("cds_conv3d_k3_roll", ptr(xc), ptr(packed), ptr(bc), 2, cin, cout, D, H, Wd, 1, ptr(out))
        torch.cuda.synchronize()
        close(from_blocked(out.float().cpu()), ref, 6e-3, 2e-3)


@pytest.mark.parametrize("cin,cout", [(64, 32), (32, 16)])
@pytest.mark.parametrize("shape", [(3, 5, 7), (2, 4, 130), (6, 37, 50)])
@pytest.mark.parametrize("with_skip", [True, False])
def test_deconv3d_gather_tcgen05_vs_torch(cin, cout, shape, with_skip):
    """Gather-form tcgen05 transposed conv (one output-parity class per blockIdx.y) against the published operator."""
    D, H, Wd = shape
    torch.manual_seed(cin + Wd)
    x = torch.randn(2, cin, D, H, Wd).half().float()
    w = torch.randn(cin, cout, 3, 3, 3) / (8 * cin) ** 0.5
    b = torch.randn(cout)
    skip = torch.randn(2, cout, 2 * D, 2 * H, 2 * Wd).half().float()
    ref = torch.relu(torch.nn.functional.conv_transpose3d(x, w, b, stride=2, padding=1, output_padding=1))
    if with_skip:
        ref = skip + ref
    lw = W.Conv3dWeights(cin, cout, w.permute(2, 3, 4, 0, 1).reshape(27, cin, cout).contiguous(), b)
    packed = cu(W.pack_deconv3d_gtc(lw))
    lib = _lib.LIB.load()
    assert lib.cds_deconv3d_k3s2_gtc_supported(cin, cout) == 1
    assert packed.numel() == lib.cds_deconv3d_k3s2_gtc_weight_halfs(cin, cout)
    xc, sc, bc = cu(to_blocked(x)).half(), cu(to_blocked(skip)).half(), cu(b)
    out = torch.full_like(sc, float("nan"))
    call("cds_deconv3d_k3s2_gtc", ptr(xc), ptr(packed), ptr(bc), ptr(sc) if with_skip else None, 2, cin, cout, D, H, Wd, ptr(out))
    torch.cuda.synchronize()
    close(from_blocked(out.float().cpu()), ref, 8e-3, 2e-3)


def _in_norm_lrelu(x, eps=1e-5):
    """InstanceNorm2d(affine=False) + LeakyReLU(0.1) of an NHWC tensor, and the (sum, sumsq) statistics the kernels consume."""
    m = x.mean(dim=(1, 2), keepdim=True)
    v = x.var(dim=(1, 2), unbiased=False, keepdim=True)
    y = torch.nn.functional.leaky_relu((x - m) / torch.sqrt(v + eps), 0.1)
    stats = torch.stack((x.double().sum(dim=(1, 2)), (x.double() ** 2).sum(dim=(1, 2))), dim=-1)   # [n, C, 2]
    return y, stats.contiguous()


@pytest.mark.parametrize("cin,cout", [(8, 16), (16, 32)])
@pytest.mark.parametrize("hw", [(32, 40), (37, 51), (64, 130)])
def test_conv2d_3x3s2_tcgen05_vs_torch(cin, cout, hw):
    """FeatureNet.downsample1/2 on the tensor cores: InstanceNorm + LeakyReLU applied on load, 3x3 stride-2 conv, raw output +
    statistics, against torch on the same fp16-stored input (odd sizes: ceil(n/2) outputs, partial tiles, borders)."""
    H, Wd = hw
    torch.manual_seed(cin + H)
    x = (torch.randn(3, H, Wd, cin) * 2 + 0.5).half().float()
    w = torch.randn(cout, cin, 3, 3) / (9 * cin) ** 0.5
    xn, stats = _in_norm_lrelu(x)
    ref = torch.nn.functional.conv2d(xn.permute(0, 3, 1, 2), w, stride=2, padding=1).permute(0, 2, 3, 1).contiguous()
    packed = cu(W.pack_conv2d_gtc(w.permute(2, 3, 1, 0).reshape(9, cin, cout)))
    lib = _lib.LIB.load()
    assert lib.cds_conv2d_3x3s2_tc_supported(cin, cout) == 1
    assert packed.numel() == lib.cds_conv2d_3x3s2_tc_weight_halfs(cin, cout)
    out = torch.full(tuple(ref.shape), float("nan"), device=DEV, dtype=torch.float16)
    ostats = torch.zeros(3, cout, 2, device=DEV, dtype=torch.float64)
    xc, sc = cu(x).half(), cu(stats)     # named: a temporary inside the argument list would be freed before the launch
    out_lo = torch.full_like(out, float("nan"))
    call("cds_conv2d_3x3s2_tc", ptr(xc), None, ptr(sc), _lib.ACT_LRELU, ptr(packed), 3, cin, cout, H, Wd, ptr(out), ptr(out_lo), ptr(ostats))
    torch.cuda.synchronize()
    # value + residual planes together carry the fp32 result to ~22 bits
    close(out.float().cpu() + out_lo.float().cpu(), ref, 3e-4, 1e-4)
    # operands are rounded to fp16 after normalisation (|x| <~ 4 -> 1e-3 abs), 9*cin terms
    close(out.float().cpu(), ref, 6e-3, 2e-3)
    ref_stats = torch.stack((ref.double().sum(dim=(1, 2)), (ref.double() ** 2).sum(dim=(1, 2))), dim=-1)
    torch.testing.assert_close(ostats.cpu(), ref_stats, atol=0.02 * ref[0, :, :, 0].numel() ** 0.5, rtol=5e-3)


@pytest.mark.parametrize("cin,cout", [(8, 16), (16, 32)])
@pytest.mark.parametrize("hw", [(32, 40), (37, 50), (64, 300), (150, 700), (9, 2)], ids=lambda hw: f"{hw[0]}x{hw[1]}")
@pytest.mark.parametrize("use_lo", [1, 0])
def test_conv2d_3x3s2_rows_vs_torch(cin, cout, hw, use_lo):
    """FeatureNet.downsample1/2 as the row-streaming kernel (csrc/conv2d_s2rows.cu): even / odd pixel phases by TMA, each input
    row normalised once, against torch on the fp32 input (value + residual planes) or its fp16 rounding (value plane only);
    odd heights, partial strips, several tiles per persistent CTA, statistics."""
    H, Wd = hw
    torch.manual_seed(cin + H + Wd)
    n = 3
    x = 0.5 + 1.3 * torch.randn(n, cin, H, Wd)
    wt = torch.randn(cout, cin, 3, 3) / (9 * cin) ** 0.5
    nhwc = x.permute(0, 2, 3, 1).contiguous()
    hi = nhwc.half(); lo = (nhwc - hi.float()).half()
    stored = (hi.float() + lo.float()) if use_lo else hi.float()
    xs = stored.permute(0, 3, 1, 2)
    ref = torch.nn.functional.conv2d(torch.nn.functional.leaky_relu(O.instance_norm(xs), 0.1), wt, stride=2, padding=1)
    stats = cu(torch.stack((xs.double().sum((2, 3)), (xs.double() ** 2).sum((2, 3))), -1).contiguous())
    lib = _lib.LIB.load()
    assert lib.cds_conv2d_3x3s2_rows_supported(cin, cout, H, Wd) == 1
    packed = cu(W.pack_conv2d_s2rows(wt.permute(2, 3, 1, 0).reshape(9, cin, cout).contiguous()))
    assert packed.numel() == lib.cds_conv2d_3x3s2_rows_weight_halfs(cin, cout)
    Ho, Wo = (H + 1) // 2, Wd // 2
    out = torch.full((n, Ho, Wo, cout), float("nan"), device=DEV, dtype=torch.float16)
    out_lo = torch.full_like(out, float("nan"))
    ostats = torch.zeros(n, cout, 2, device=DEV, dtype=torch.float64)
    hc, lc = cu(hi), cu(lo)
    call("cds_conv2d_3x3s2_rows", ptr(hc), ptr(lc) if use_lo else None, ptr(stats), _lib.ACT_LRELU, ptr(packed), n, cin, cout, H, Wd,
         ptr(out), ptr(out_lo), ptr(ostats))
    torch.cuda.synchronize()
    y = (out.float() + out_lo.float()).cpu().permute(0, 3, 1, 2)
    assert not torch.isnan(y).any()
    err = O.rel_l1(y, ref)
    print(f"3x3 s2 rows {cin}->{cout} {hw} lo={use_lo}: {err:.2e}")
    assert err < 2e-5
    assert (out_lo.float().abs() <= out.float().abs() * 2.0 ** -11 + 1e-7).all()
    ref_stats = torch.stack((y.double().sum(dim=(2, 3)), (y.double() ** 2).sum(dim=(2, 3))), dim=-1)
    torch.testing.assert_close(ostats.cpu(), ref_stats, atol=0.02 * (Ho * Wo) ** 0.5, rtol=2e-3)
    # and against the gather form of the same layer
    packed_g = cu(W.pack_conv2d_gtc(wt.permute(2, 3, 1, 0).reshape(9, cin, cout).contiguous()))
    out_g = torch.empty_like(out); out_g_lo = torch.empty_like(out)
    call("cds_conv2d_3x3s2_tc", ptr(hc), ptr(lc) if use_lo else None, ptr(stats), _lib.ACT_LRELU, ptr(packed_g), n, cin, cout, H, Wd,
         ptr(out_g), ptr(out_g_lo), None)
    torch.cuda.synchronize()
    assert O.rel_l1(y, (out_g.float() + out_g_lo.float()).cpu().permute(0, 3, 1, 2)) < 2e-6


def test_conv2d_3x3s2_rows_rejects_odd_width():
    assert _lib.LIB.load().cds_conv2d_3x3s2_rows_supported(8, 16, 32, 41) == 0
    assert _lib.LIB.load().cds_conv2d_3x3s2_rows_supported(8, 8, 32, 40) == 0


@pytest.mark.parametrize("ca,cb,cout,a_norm", [(32, 16, 16, True), (16, 8, 8, False)])
@pytest.mark.parametrize("hw", [(32, 40), (38, 50), (64, 130)])
def test_conv2d_1x1_cat_tcgen05_vs_torch(ca, cb, cout, a_norm, hw):
    """FeatureNet.inner1/2 on the tensor cores: 1x1 conv over cat(nearest-up2(a), b), both normalised on load (inner2 takes
    the already-activated stage-2 feature as is)."""
    H, Wd = hw
    torch.manual_seed(ca + H)
    a = (torch.randn(2, H // 2, Wd // 2, ca) * (2 if a_norm else 0.5)).half().float()
    b = (torch.randn(2, H, Wd, cb) * 3 - 1).half().float()
    w = torch.randn(ca + cb, cout) / (ca + cb) ** 0.5
    an, a_stats = _in_norm_lrelu(a) if a_norm else (a, None)
    bn, b_stats = _in_norm_lrelu(b)
    up = an.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    ref = torch.cat((up, bn), dim=-1) @ w
    packed = cu(W.pack_conv2d_gtc(w.reshape(1, ca + cb, cout)))
    lib = _lib.LIB.load()
    assert lib.cds_conv2d_1x1_cat_tc_supported(ca, cb, cout) == 1
    assert packed.numel() == lib.cds_conv2d_1x1_cat_tc_weight_halfs(ca, cb, cout)
    out = torch.full(tuple(ref.shape), float("nan"), device=DEV, dtype=torch.float16)
    ostats = torch.zeros(2, cout, 2, device=DEV, dtype=torch.float64)
    ac, bc, asc, bsc = cu(a).half(), cu(b).half(), (cu(a_stats) if a_norm else None), cu(b_stats)
    call("cds_conv2d_1x1_cat_tc", ptr(ac), ptr(asc), _lib.ACT_LRELU if a_norm else _lib.ACT_NONE,
         ptr(bc), ptr(bsc), _lib.ACT_LRELU, ptr(packed), 2, ca, cb, cout, H, Wd, ptr(out), ptr(ostats))
    torch.cuda.synchronize()
    close(out.float().cpu(), ref, 6e-3, 2e-3)
    ref_stats = torch.stack((ref.double().sum(dim=(1, 2)), (ref.double() ** 2).sum(dim=(1, 2))), dim=-1)
    torch.testing.assert_close(ostats.cpu(), ref_stats, atol=0.02 * ref[0, :, :, 0].numel() ** 0.5, rtol=5e-3)


@pytest.mark.parametrize("st", [0, 1, 2])
@pytest.mark.parametrize("hw", [(37, 200), (8, 128), (21, 300), (32, 40), (19, 100)])
def test_visnet_tcgen05_vs_oracle(pretrained_sd, st, hw):
    """tcgen05 visibility net (3 chained tap-GEMM layers) against the oracle."""
    torch.manual_seed(st * 10 + hw[0])
    x = torch.cat((3.5 * torch.rand(3, 1, *hw), 0.3 * torch.rand(3, 1, *hw)), 1)
    ref = O.vis_net(x, pretrained_sd, f"stage_net.vis.{st}")
    wgt, fp = W.pack_visnet_tc(pretrained_sd, f"stage_net.vis.{st}", DEV)
    assert wgt.numel() == _lib.LIB.load().cds_visnet_tc_weight_halfs()
    xc = cu(x)
    ent, cur = xc[:, 0].contiguous(), xc[:, 1].contiguous()
    out = torch.full((3, *hw), -1.0, device=DEV)
    call("cds_visnet_tc", ptr(ent), ptr(cur), ptr(wgt), ptr(fp), 3, hw[0], hw[1], ptr(out))
    torch.cuda.synchronize()
    assert out.min() >= 0            # every pixel written exactly by its owner tile
    close(out.unsqueeze(1), ref, 4e-3, 1e-3)


@pytest.mark.parametrize("layer", ["conv01", "conv10", "conv20"])
@pytest.mark.parametrize("hw", [(24, 40), (37, 150)])
def test_dynamic_conv_tcgen05_split_precision(pretrained_sd, hw, layer):
    """Split-precision input (fp16 value + fp16 residual planes) with the producer's InstanceNorm + LeakyReLU applied on
    load, for every trunk layer shape (8->8 (3,5,7), 16->16 (3,5), 32->32 (1,3)): must be clearly more accurate than the
    single-plane path, and both must match the oracle."""
    import ctypes
    cin, cout, ks, pre = W.DYN_LAYERS[layer]
    torch.manual_seed(hw[1])
    n = 2
    x = 1.7 + 0.8 * torch.randn(n, cin, *hw)                       # raw pre-norm activations with a mean offset
    epi = torch.tensor([[hw[1] * 1.7, -hw[0] * 0.6], [-30.0, hw[0] / 2.0]])
    xin = torch.nn.functional.leaky_relu(O.instance_norm(x), 0.1)
    ref_y, ref_nc = O.dynamic_conv(xin, pretrained_sd, pre, ks, epi, T)
    w = W.pack_dynamic_conv(pretrained_sd, pre, cin, cout, ks, DEV)
    w.tc = W.pack_dynamic_conv_tc(w)
    nhwc = x.permute(0, 2, 3, 1).contiguous()
    hi = nhwc.half()
    lo = (nhwc - hi.float()).half()
    stats = torch.stack((x.double().sum((2, 3)), (x.double() ** 2).sum((2, 3))), -1).contiguous()   # [n, C, 2]
    kz = (ctypes.c_int * len(ks))(*ks)
    errs = {}
    for split in (1, 0):
        planes = cu(torch.stack((hi, lo)) if split else hi.unsqueeze(0)).contiguous()
        st = cu(stats)
        if not split:   # statistics of what is actually stored
            xs = hi.float().permute(0, 3, 1, 2).double()
            st = cu(torch.stack((xs.sum((2, 3)), (xs ** 2).sum((2, 3))), -1).contiguous())
        out = torch.empty(n, *hw, cout, device=DEV, dtype=torch.float16)
        out_lo = torch.empty_like(out)
        nc = torch.empty(n, *hw, device=DEV)
        epi_c = cu(epi)
        call("cds_dynamic_conv_tc", ptr(planes), n, None, ptr(st), 1, ptr(epi_c), 1.0, ptr(w.tc), None, ptr(w.gate), n, cin, cout,
             hw[0], hw[1], len(ks), kz, T, split, ptr(out), ptr(out_lo), None, ptr(nc), None, 0, None)
        torch.cuda.synchronize()
        y = (out.float() + out_lo.float()).cpu().permute(0, 3, 1, 2)
        errs[split] = (O.rel_l1(y, ref_y), O.rel_l1(nc.cpu().unsqueeze(1), ref_nc))
    print(f"split-precision {layer} rel-L1 (out, curv): split", errs[1], " single", errs[0])
    assert errs[0][0] < 4e-3 and errs[0][1] < 4e-3
    assert errs[1][0] < 1e-3 and errs[1][1] < 1e-3
    assert errs[1][0] < 0.7 * errs[0][0]


def test_dynamic_conv_pairs_residual_plane(pretrained_sd):
    """conv00 over the pair batch also emits the rounding-residual plane of its output: value + residual is what the fp32
    accumulators held (error of the pair far below one fp16 ulp of the value plane alone)."""
    import ctypes
    cin, cout, ks, pre = W.DYN_LAYERS["conv00"]
    V, B, hw = 2, 1, (40, 150)
    torch.manual_seed(3)
    imgs = torch.rand(B * (V + 1), 3, *hw)
    w = W.pack_dynamic_conv(pretrained_sd, pre, cin, cout, ks, DEV)
    w.tc = W.pack_dynamic_conv_tc(w)
    epi = torch.tensor([[200.0, -30.0], [-50.0, 20.0], [10.0, 300.0], [400.0, 100.0]])   # (side, v, b)
    idx = torch.tensor([0, 0, 1, 2], dtype=torch.int32)
    img8 = torch.empty(B * (V + 1), *hw, 8, device=DEV, dtype=torch.float16)
    ic = cu(imgs)
    call("cds_image_to_nhwc8", ptr(ic), B * (V + 1), hw[0], hw[1], ptr(img8))
    n = 2 * V * B
    out = torch.empty(n, *hw, cout, device=DEV, dtype=torch.float16)
    out_lo = torch.empty_like(out)
    kz = (ctypes.c_int * 3)(*ks)
    ec, xc = cu(epi), cu(idx)
    call("cds_dynamic_conv_tc_pairs", ptr(img8), B * (V + 1), ptr(xc), ptr(ec), 1.0, ptr(w.tc), None, ptr(w.gate), V, B, 8, cout,
         hw[0], hw[1], 3, kz, T, ptr(out), ptr(out_lo), None, None, None, 0, None)
    torch.cuda.synchronize()
    for i in range(n):
        ref_y, _ = O.dynamic_conv(imgs[int(idx[i])][None], pretrained_sd, pre, ks, epi[i][None], T)
        one = O.rel_l1(out[i].float().cpu().permute(2, 0, 1)[None], ref_y)
        two = O.rel_l1((out[i].float() + out_lo[i].float()).cpu().permute(2, 0, 1)[None], ref_y)
        print(f"conv00 item {i}: value plane {one:.2e}, value + residual {two:.2e}")
        assert two < 5e-5 and two < 0.5 * one


def test_conv2d_3x3s2_reads_residual_plane():
    """downsample1/2 fed value + residual planes equal the conv of the fp32 input (not of its fp16 rounding)."""
    torch.manual_seed(5)
    n, cin, cout, hw = 2, 8, 16, (38, 70)
    x = 0.5 + 1.3 * torch.randn(n, cin, *hw)
    wt = torch.randn(cout, cin, 3, 3) * 0.2
    ref = torch.nn.functional.conv2d(torch.nn.functional.leaky_relu(O.instance_norm(x), 0.1), wt, stride=2, padding=1)
    nhwc = x.permute(0, 2, 3, 1).contiguous()
    hi = nhwc.half(); lo = (nhwc - hi.float()).half()
    stats = cu(torch.stack((x.double().sum((2, 3)), (x.double() ** 2).sum((2, 3))), -1).contiguous())
    packed = cu(W.pack_conv2d_gtc(wt.permute(2, 3, 1, 0).reshape(9, cin, cout).contiguous()))
    Ho, Wo = (hw[0] + 1) // 2, (hw[1] + 1) // 2
    errs = []
    for use_lo in (True, False):
        out = torch.empty(n, Ho, Wo, cout, device=DEV, dtype=torch.float16)
        out_lo = torch.empty_like(out)
        hc, lc = cu(hi), cu(lo)
        call("cds_conv2d_3x3s2_tc", ptr(hc), ptr(lc) if use_lo else None, ptr(stats), _lib.ACT_LRELU, ptr(packed), n, cin, cout, hw[0], hw[1],
             ptr(out), ptr(out_lo), None)
        torch.cuda.synchronize()
        errs.append(O.rel_l1((out.float() + out_lo.float()).cpu().permute(0, 3, 1, 2), ref))
    print("3x3 s2 conv with / without the input's residual plane:", errs)
    assert errs[0] < 2e-5 and errs[0] < 0.3 * errs[1]


def test_instnorm_act_split_f32():
    torch.manual_seed(1)
    n, C_, hw = 3, 32, (20, 36)
    x = 2.0 + torch.randn(n, C_, *hw)
    ref = torch.tanh(O.instance_norm(x))
    nhwc = x.permute(0, 2, 3, 1).contiguous()
    hi = nhwc.half(); lo = (nhwc - hi.float()).half()
    stats = cu(torch.stack((x.double().sum((2, 3)), (x.double() ** 2).sum((2, 3))), -1).contiguous())
    out = torch.empty(n, *hw, C_, device=DEV)
    hc, lc = cu(hi), cu(lo)
    out16 = torch.empty(n, *hw, C_, device=DEV, dtype=torch.float16)
    call("cds_instnorm_act_split_f32", ptr(hc), ptr(lc), ptr(stats), _lib.ACT_TANH, n, C_, hw[0], hw[1], ptr(out), ptr(out16))
    torch.cuda.synchronize()
    assert O.rel_l1(out.cpu().permute(0, 3, 1, 2), ref) < 2e-6
    assert torch.equal(out16, out.half())


@pytest.mark.parametrize("shape", [(2, 1, 32, 8, 24, 40), (1, 2, 32, 5, 17, 23), (3, 1, 32, 6, 31, 45), (4, 2, 32, 4, 16, 33),
                                   (5, 1, 32, 4, 16, 24), (2, 1, 16, 8, 24, 40)], ids=lambda v: "x".join(map(str, v)))
def test_costvol_aggregate_split_matches_fp32_form(shape):
    """fp32 features -> split fp16 volume: value + residual planes reproduce the fp32-storage aggregate.  C = 32 with up to
    4 views runs the four-channels-per-lane sweep, the fp32-storage call the eight-channel one: bit-identical volumes."""
    torch.manual_seed(2)
    V, B, C_, D, h, w = shape
    ref_f = cu(torch.tanh(torch.randn(V, B, h, w, C_))); src_f = cu(torch.tanh(torch.randn(V, B, h, w, C_)))
    s = synthetic.make_sample(dict(W=4 * w, H=4 * h, N=V + 1, ndepths=(D,), ratios=(1.0,), B=B, Dtot=192, interval=2.65))
    pm = cu(s.proj_matrices["stage1"])
    coef = torch.empty(1, B, V, 12, device=DEV)
    import ctypes
    arr = (ctypes.c_void_p * 1)(pm.data_ptr())
    call("cds_camera_setup", arr, 1, 0, B, V + 1, ptr(coef), None)
    dv = cu((500 + 20 * torch.arange(D).float()).reshape(1, D, 1, 1).expand(B, D, h, w).contiguous())
    vis = cu(torch.rand(V, B, h, w))
    full = torch.empty(B, C_ // 8, D, h, w, 8, device=DEV)
    call("cds_costvol_aggregate", ptr(ref_f), ptr(src_f), ptr(coef), ptr(dv), ptr(vis), V, B, C_, D, h, w, _lib.CDS_F32, ptr(full))
    hi = torch.empty(B, C_ // 8, D, h, w, 8, device=DEV, dtype=torch.float16)
    lo = torch.empty_like(hi)
    call("cds_costvol_aggregate_split", ptr(ref_f), ptr(src_f), ptr(coef), ptr(dv), ptr(vis), V, B, C_, D, h, w, ptr(hi), ptr(lo))
    torch.cuda.synchronize()
    assert torch.equal(hi, full.half())
    assert (hi.float() + lo.float() - full).abs().max() < 1e-6 * full.abs().max() + 1e-7
