"""Parity of the production path (fp16 storage, tcgen05 kernels, CUDA graph) at north_star's bar -- depth within 1e-3
relative L1 of the reference on identical inputs -- where round 1 had not proven it:

* every tested seed of the worst-case "noise" input family (uniform-noise images: no photo-consistent surface, the softmaxes
  with 1/T = 100 amplify rounding), not just the committed golden's seed 0;
* BASELINE.json's full-size configurations against the CPU oracle (cfg2 noise + plane, cfg4 with B = 4);
* against the LIVE reference (oracle/_ref, the unmodified ``models`` package run eagerly on the same GPU with TF32 off) at
  cfg2 / cfg3 / cfg5, and the reference's OWN ``CDSMVSNet`` / ``StageNet`` driver code running on the CUDA operators through
  ``patch()`` at every level.
"""
import numpy as np
import pytest
import torch

import cds_mvsnet_b200 as C
from cds_mvsnet_b200 import synthetic
from oracle import oracle as O
from oracle import ref_live

pytestmark = pytest.mark.gpu
T = 0.01
DEV = "cuda"
DEPTH_REL_L1 = 1e-3   # north_star tolerance
torch.set_grad_enabled(False)


def build(sd, cfg, storage=torch.float16, refine=False):
    m = C.CDSMVSNet(refine=refine, ndepths=cfg["ndepths"], depth_interals_ratio=cfg["ratios"], storage=storage)
    m.load_state_dict({k: v for k, v in sd.items() if k in m.state_dict()}, strict=True)
    return m.to(DEV).eval()


def run(model, s):
    return model(s.imgs.to(DEV), {k: v.to(DEV) for k, v in s.proj_matrices.items()}, s.depth_values.to(DEV), temperature=T)


def check_stages(tag, out, ref, n_stages, tol=DEPTH_REL_L1, conf_tol=8e-2):
    """Depth: north_star's bar.  photometric_confidence is a HARD window pick (sum of the 4 probabilities around
    trunc(expected index), models/module.py:382-391): where the expected index sits near an integer a 1e-4 change of depth
    flips the window, so its mean error is bounded loosely (measured 4.6e-2 on the flat distributions of the noise family at
    cfg2's stage 3, 3e-4 on photo-consistent input)."""
    worst = 0.0
    for st in range(1, n_stages + 1):
        d, r = out[f"stage{st}"]["depth"].float().cpu(), ref[f"stage{st}"]["depth"].float().cpu()
        assert d.shape == r.shape
        rel = O.rel_l1(d, r)
        cerr = (out[f"stage{st}"]["photometric_confidence"].float().cpu() - ref[f"stage{st}"]["photometric_confidence"].float().cpu()).abs().mean().item()
        print(f"{tag} stage{st}: depth rel-L1 {rel:.3e}  conf |err| {cerr:.3e}")
        assert rel < tol, (tag, st, rel)
        assert cerr < conf_tol, (tag, st, cerr)
        worst = max(worst, rel)
    return worst


SEED_CFG = dict(W=160, H=128, N=4, ndepths=(48, 32, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=192, interval=2.65)


@pytest.mark.parametrize("seed", range(6))
def test_noise_seeds_vs_oracle(pretrained_sd, seed):
    """the default path meets 1e-3 on EVERY tested seed of the chaotic input, per stage."""
    s = synthetic.make_sample(SEED_CFG, "noise", seed=seed)
    O.FAST_GATHER = True
    ref = O.cdsmvsnet_forward(pretrained_sd, s.imgs, s.proj_matrices, s.depth_values, SEED_CFG["ndepths"], SEED_CFG["ratios"], T)
    out = run(build(pretrained_sd, SEED_CFG), s)
    check_stages(f"noise seed {seed}", out, ref, 3)


@pytest.mark.parametrize("name,family", [("cfg2", "noise"), ("cfg2", "plane"), ("cfg4", "noise")])
def test_full_size_vs_oracle(pretrained_sd, name, family):
    """The headline configuration (1600x1184, N=5, D=48/32/8) and the B=4 training shape against the CPU oracle itself."""
    cfg = synthetic.CONFIGS[name]
    s = synthetic.make_sample(name, family, seed=0)
    O.FAST_GATHER = True
    torch.set_num_threads(max(1, (torch.get_num_threads())))
    ref = O.cdsmvsnet_forward(pretrained_sd, s.imgs, s.proj_matrices, s.depth_values, cfg["ndepths"], cfg["ratios"], T)
    out = run(build(pretrained_sd, cfg), s)
    check_stages(f"{name} {family}", out, ref, len(cfg["ndepths"]))


def _live(sd, cfg, s, refine=False):
    """The unmodified reference, eager on the GPU, fp32 (TF32 off -- its GPU default would round to 10 bits)."""
    tf = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        m = ref_live.build_model(sd, cfg["ndepths"], cfg["ratios"], refine=refine, device=DEV)
        out = run(m, s)
        out = {k: ({kk: vv.float().cpu() for kk, vv in v.items()} if isinstance(v, dict) else v.float().cpu()) for k, v in out.items()}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf
    del m
    torch.cuda.empty_cache()
    return out


needs_ref = pytest.mark.skipif(not ref_live.available(), reason="oracle/_ref/reference_models.zip not shipped (run build())")


@needs_ref
@pytest.mark.parametrize("name,family", [("cfg2", "plane"), ("cfg2", "noise"), ("cfg3", "plane"), ("cfg5", "plane"), ("cfg4", "plane")])
def test_full_size_vs_live_reference(pretrained_sd, name, family):
    """Every BASELINE.json configuration against the LIVE reference run on the same GPU."""
    cfg = synthetic.CONFIGS[name]
    s = synthetic.make_sample(name, family, seed=1)
    ref = _live(pretrained_sd, cfg, s)
    out = run(build(pretrained_sd, cfg), s)
    check_stages(f"live {name} {family}", out, ref, len(cfg["ndepths"]))


@needs_ref
@pytest.mark.parametrize("level", C.PATCH_LEVELS)
@pytest.mark.parametrize("family", ["plane", "noise"])
def test_patch_drops_into_live_reference(pretrained_sd, level, family):
    """north_star's call-surface claim, executed: ``patch(models.model, models.module, level)`` and then the reference's own
    constructor + forward (models/model.py:97-223) -- at "leaf"/"ops" its own CDSMVSNet.forward and StageNet.forward drive the
    CUDA operators, at "stage"/"model" the fused ones -- against the unpatched reference on the same input."""
    cfg = dict(W=160, H=128, N=3, ndepths=(16, 8, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=192, interval=2.65)
    s = synthetic.make_sample(cfg, family, seed=0)
    ref = _live(pretrained_sd, cfg, s)
    rmodel, rmodule, _, _ = ref_live.load()
    saved = C.patch(rmodel, rmodule, level=level)
    try:
        m = ref_live.build_model(pretrained_sd, cfg["ndepths"], cfg["ratios"], device=DEV, rmodel=rmodel)
        if level == "model":
            assert isinstance(m, C.CDSMVSNet)
        else:
            assert type(m).__module__ == "models.model" and isinstance(m.cost_regularization[0], C.CostRegNet)
            assert isinstance(m.feature, C.FeatureNet) == (level != "leaf")
            assert isinstance(m.stage_net, C.StageNet) == (level == "stage")
        out = run(m, s)
    finally:
        C.unpatch(saved)
    assert rmodel.CDSMVSNet.__module__ == "models.model" and rmodule.DynamicConv.__module__ == "models.dynamic_conv"
    check_stages(f"patch[{level}] {family}", out, ref, 3)
    assert set(out) >= {"stage1", "stage2", "stage3", "depth", "photometric_confidence", "refined_depth"}


@needs_ref
def test_patch_refine_true_live(pretrained_sd, golden):
    """refine=True (the configuration of every pretrained checkpoint) through the patched reference constructor."""
    sd = dict(pretrained_sd)
    sd.update(golden("weights_refine_both_dtu_blended"))
    cfg = dict(W=256, H=192, N=3, ndepths=(16, 8, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=192, interval=2.65, refine=True)
    s = synthetic.make_sample(cfg, "plane", seed=0)
    ref = _live(sd, cfg, s, refine=True)
    rmodel, rmodule, _, _ = ref_live.load()
    saved = C.patch(rmodel, rmodule, level="model")
    try:
        m = ref_live.build_model(sd, cfg["ndepths"], cfg["ratios"], refine=True, device=DEV, rmodel=rmodel)
        out = run(m, s)
    finally:
        C.unpatch(saved)
    check_stages("patch[model] refine", out, ref, 3)
    assert O.rel_l1(out["refined_depth"].float().cpu(), ref["refined_depth"]) < DEPTH_REL_L1


def test_uint8_images_match_float_images(pretrained_sd):
    """8-bit images uploaded as bytes and divided by 255 on the device give the maps of the float images bit for bit up to the
    atomics' re-association (the IEEE quotient is the host's)."""
    cfg = dict(W=160, H=128, N=3, ndepths=(16, 8, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=192, interval=2.65)
    s = synthetic.make_sample(cfg, "plane", seed=3)
    u8 = (s.imgs * 255.0).round().clamp(0, 255).to(torch.uint8)
    f32 = torch.from_numpy(u8.numpy().astype(np.float32) / np.float32(255.0))
    model = build(pretrained_sd, cfg)
    proj = {k: v.to(DEV) for k, v in s.proj_matrices.items()}
    a = model(f32.to(DEV), proj, s.depth_values.to(DEV), temperature=T)
    b = model(u8.to(DEV), proj, s.depth_values.to(DEV), temperature=T)
    for st in (1, 2, 3):
        assert O.rel_l1(b[f"stage{st}"]["depth"].cpu(), a[f"stage{st}"]["depth"].cpu()) < 2e-5


def test_stale_weight_cache_is_detected(pretrained_sd):
    """reloading a SUBMODULE or editing a parameter in place must not leave the folded weights stale."""
    cfg = dict(W=160, H=128, N=3, ndepths=(16, 8, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=192, interval=2.65)
    s = synthetic.make_sample(cfg, "plane", seed=1)
    model = build(pretrained_sd, cfg)
    a = run(model, s)["depth"].clone()
    fsd = {k: v.clone() for k, v in model.feature.state_dict().items()}
    key = "conv01.conv.convs.1.weight"
    fsd[key] = fsd[key] * 1.5
    model.feature.load_state_dict(fsd)                       # a submodule's load: the parent's hooks do not fire
    b = run(model, s)["depth"].clone()
    assert O.rel_l1(b.cpu(), a.cpu()) > 1e-6, "folded weights were not rebuilt after a submodule load_state_dict"
    with torch.no_grad():
        model.feature.conv01.conv.convs[1].weight.div_(1.5)   # in-place edit (what an optimizer step / EMA swap does)
    c = run(model, s)["depth"]
    assert O.rel_l1(c.cpu(), a.cpu()) < 2e-5, "folded weights were not rebuilt after an in-place parameter edit"
    # an edit through ``.data`` bypasses autograd's version counter: the documented escape hatch is invalidate_cache()
    model.feature.conv01.conv.convs[1].weight.data.mul_(1.5)
    model.invalidate_cache()
    d = run(model, s)["depth"]
    assert O.rel_l1(d.cpu(), b.cpu()) < 2e-5


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_model_on_second_device(pretrained_sd):
    """a model on cuda:1 while cuda:0 is current launches on cuda:1 with cuda:1's stream."""
    cfg = dict(W=160, H=128, N=3, ndepths=(16, 8, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=192, interval=2.65)
    s = synthetic.make_sample(cfg, "plane", seed=1)
    a = run(build(pretrained_sd, cfg), s)["depth"].cpu()
    torch.cuda.set_device(0)
    m = C.CDSMVSNet(ndepths=cfg["ndepths"], depth_interals_ratio=cfg["ratios"])
    m.load_state_dict(pretrained_sd)
    m = m.to("cuda:1").eval()
    out = m(s.imgs.to("cuda:1"), {k: v.to("cuda:1") for k, v in s.proj_matrices.items()}, s.depth_values.to("cuda:1"), temperature=T)
    assert out["depth"].device.index == 1 and torch.cuda.current_device() == 0
    assert O.rel_l1(out["depth"].cpu(), a) < 2e-5
