"""End-to-end parity of the fused cascade (drop-in CDSMVSNet) against golden outputs of the live reference
and against the CPU oracle.  The bar is north_star's: depth within 1e-3 relative L1 of the reference."""
import os

import pytest
import torch

import cds_mvsnet_b200 as C
from cds_mvsnet_b200 import synthetic
from oracle import oracle as O

pytestmark = pytest.mark.gpu
T = 0.01
DEV = "cuda"
DEPTH_REL_L1 = 1e-3      # north_star tolerance (fp16 storage, fp32 accumulate)
DEPTH_REL_L1_F32 = 2e-5  # fp32-storage build: re-association noise only
torch.set_grad_enabled(False)


def build(sd, nd, ratios, storage):
    m = C.CDSMVSNet(refine=False, ndepths=nd, depth_interals_ratio=ratios, storage=storage)
    m.load_state_dict({k: v for k, v in sd.items() if k in m.state_dict()}, strict=True)
    return m.to(DEV).eval()


def cfg_of(g):
    W, H, N, B, Dtot = (int(v) for v in g["cfg"][:5])
    nd = tuple(int(v) for v in g["cfg"][5:])
    return dict(W=W, H=H, N=N, B=B, Dtot=Dtot, ndepths=nd, ratios=tuple(float(r) for r in g["ratios"]),
                interval=float(g["interval"]))


def run(model, s):
    return model(s.imgs.to(DEV), {k: v.to(DEV) for k, v in s.proj_matrices.items()}, s.depth_values.to(DEV), temperature=T)


@pytest.mark.parametrize("storage", [torch.float32, torch.float16])
@pytest.mark.parametrize("tag,family", [("e2e_cfg1_noise", "noise"), ("e2e_small3_plane", "plane"), ("e2e_small3_noise", "noise")])
def test_cascade_golden(golden, pretrained_sd, tag, family, storage):
    g = golden(tag)
    cfg = cfg_of(g)
    s = synthetic.make_sample(cfg, family, seed=0)
    model = build(pretrained_sd, cfg["ndepths"], cfg["ratios"], storage)
    out = run(model, s)
    assert set(out) >= {"depth", "photometric_confidence", "norm_curv", "refined_depth", "stage1"}
    tol = DEPTH_REL_L1_F32 if storage == torch.float32 else DEPTH_REL_L1
    for st in range(len(cfg["ndepths"])):
        nm = f"stage{st + 1}"
        ref_d, ref_c, ref_n = g[f"{nm}_depth"], g[f"{nm}_photometric_confidence"], g[f"{nm}_norm_curv"]
        d = out[nm]["depth"].cpu()
        assert d.shape == ref_d.shape
        rel = O.rel_l1(d, ref_d)
        conf_err = (out[nm]["photometric_confidence"].cpu() - ref_c).abs().mean().item()
        print(f"{tag} {nm} storage={storage}: depth rel-L1 {rel:.3e}  conf |err| {conf_err:.3e}")
        assert rel < tol, (nm, rel)
        assert conf_err < (1e-3 if storage == torch.float32 else 4e-2)   # confidence is a hard window index
        assert O.rel_l1(out[nm]["norm_curv"].cpu(), ref_n) < (1e-4 if storage == torch.float32 else 1e-2)
    assert torch.equal(out["refined_depth"], out["depth"])
    if family == "plane":
        assert (out["depth"].cpu() - s.gt_depth).abs().mean() < 8.0     # known answer: the plane is recovered


def test_cascade_random_weights_vs_oracle():
    """Weights the golden files do not cover: random init with randomised BN statistics, oracle computed here."""
    cfg = dict(W=128, H=96, N=3, ndepths=(16, 8, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=192, interval=2.65)
    sd = synthetic.random_state_dict(cfg["ndepths"])
    s = synthetic.make_sample(cfg, "plane", seed=2)
    ref = O.cdsmvsnet_forward(sd, s.imgs, s.proj_matrices, s.depth_values, cfg["ndepths"], cfg["ratios"], T)
    out = run(build(sd, cfg["ndepths"], cfg["ratios"], torch.float32), s)
    for st in (1, 2, 3):
        assert O.rel_l1(out[f"stage{st}"]["depth"].cpu(), ref[f"stage{st}"]["depth"]) < 1e-4
    out16 = run(build(sd, cfg["ndepths"], cfg["ratios"], torch.float16), s)
    # Untrained random weights make the depth output a chaotic function of the activations (like the "noise"
    # image family); the fp32-storage run above pins correctness, the fp16 run is only bounded loosely here.
    # The 1e-3 north_star bar is enforced on the reference's pretrained weights in test_cascade_golden.
    assert O.rel_l1(out16["depth"].cpu(), ref["depth"]) < 3e-3


def test_stagenet_dropin_vs_oracle(pretrained_sd):
    """StageNet with the reference's own signature, fed reference-layout (NCHW fp32) features."""
    torch.manual_seed(7)
    B, V, D, h, w, Cc, st = 1, 2, 8, 32, 40, 16, 1
    s = synthetic.make_sample(dict(W=2 * w, H=2 * h, N=V + 1, ndepths=(8,), ratios=(1.0,), B=B, Dtot=192, interval=2.65))
    pm = s.proj_matrices["stage3"]
    feats = [{k: (torch.tanh(torch.randn(B, Cc, h, w)), torch.rand(B, 1, h, w) * 0.01, torch.rand(B, 1, h, w) * 0.1)
              for k in ("ref", "src")} for _ in range(V)]
    dv = (425 + 60 * torch.arange(D).float()).reshape(1, D, 1, 1) + 5 * torch.rand(B, D, h, w)
    ref = O.stage_net(feats, pm, dv, pretrained_sd, st)
    net = C.StageNet(3, storage=torch.float32)
    net.load_state_dict({k[10:]: v for k, v in pretrained_sd.items() if k.startswith("stage_net.")})
    cr = C.CostRegNet(Cc, 8, storage=torch.float32)
    pre = f"cost_regularization.{st}."
    cr.load_state_dict({k[len(pre):]: v for k, v in pretrained_sd.items() if k.startswith(pre)})
    net, cr = net.to(DEV).eval(), cr.to(DEV).eval()
    feats_c = [{k: tuple(t.to(DEV) for t in v) for k, v in f.items()} for f in feats]
    out = net(feats_c, pm.to(DEV), dv.to(DEV), D, cr, stage_idx=st)
    assert set(out) == {"depth", "photometric_confidence", "norm_curv"}
    assert O.rel_l1(out["depth"].cpu(), ref["depth"]) < 2e-5
    assert O.rel_l1(out["norm_curv"].cpu(), ref["norm_curv"]) < 1e-5
    with pytest.raises(AssertionError):
        net(feats_c[:1], pm.to(DEV), dv.to(DEV), D, cr, stage_idx=st)


def test_repeatable_and_input_not_mutated(pretrained_sd):
    cfg = synthetic.CONFIGS["cfg1"]
    s = synthetic.make_sample("cfg1", "noise", seed=4)
    model = build(pretrained_sd, cfg["ndepths"], cfg["ratios"], torch.float16)
    imgs = s.imgs.to(DEV)
    keep = imgs.clone()
    a = run(model, s)["depth"]
    b = run(model, s)["depth"]
    assert torch.equal(imgs, keep)
    # InstanceNorm statistics are accumulated with fp64 atomics: run-to-run differences stay at rounding level
    assert O.rel_l1(a.cpu(), b.cpu()) < 1e-5


def _tol(name):
    """Two runs of the same kernels differ by the order of the fp64 statistics atomics (1e-16 relative), which the chaotic
    inputs amplify to ~1e-6 of depth; photometric_confidence is a HARD window pick around trunc(expected index)
    (models/module.py:382-391), so one pixel whose index sits on an integer moves its mean by 1e-5 .. 1e-4."""
    return 1e-3 if name == "photometric_confidence" else 2e-5


def test_depth_map_stream_matches_direct_call(pretrained_sd):
    """DepthMapStream (double-buffered host<->device copies on side streams) returns the same maps as the direct call for
    several different work items in flight (same kernels; the InstanceNorm statistics are accumulated with atomics, so two
    runs agree to re-association noise, not bit for bit)."""
    from cds_mvsnet_b200.streaming import DepthMapStream
    cfg = dict(W=160, H=128, N=3, ndepths=(16, 8, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=192, interval=2.65)
    model = build(pretrained_sd, cfg["ndepths"], cfg["ratios"], torch.float16)
    items = [synthetic.make_sample(cfg, "plane" if i % 2 else "noise", seed=i) for i in range(5)]
    direct = [{k: {kk: vv.cpu() for kk, vv in v.items()} for k, v in run(model, s).items() if isinstance(v, dict)} for s in items]
    stream = DepthMapStream(model, temperature=T)
    got, prev = [], None
    for s in items:
        t = stream.submit(s.imgs, s.proj_matrices, s.depth_values)
        if prev is not None:
            got.append({k: v.clone() for k, v in stream.result(prev).items()})
        prev = t
    got.append({k: v.clone() for k, v in stream.result(prev).items()})
    assert len(got) == len(items)
    for d, g in zip(direct, got):
        for st, maps in d.items():
            for name, ref in maps.items():
                err = O.rel_l1(g[f"{st}.{name}"], ref)
                assert err < _tol(name), (st, name, err)
    with pytest.raises(ValueError):
        stream.result(0)          # overwritten two submits ago
    with pytest.raises(ValueError):
        stream.submit(items[0].imgs.to(DEV), items[0].proj_matrices, items[0].depth_values)


def test_maps_in_flight_match_one_at_a_time(pretrained_sd):
    """DepthMapStream with two / three work items computing at once (one cascade clone and stream per lane) returns, item by
    item, what one map at a time returns -- including across a change of the input shape mid-stream; the clones share the
    packed weights and own their buffers."""
    from cds_mvsnet_b200.streaming import DepthMapStream
    cfg = dict(W=160, H=128, N=3, ndepths=(16, 8, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=192, interval=2.65)
    cfg_b = dict(cfg, W=192, H=96)
    model = build(pretrained_sd, cfg["ndepths"], cfg["ratios"], torch.float16)
    items = [synthetic.make_sample(cfg if i < 5 else cfg_b, "plane" if i % 2 else "noise", seed=i) for i in range(8)]
    eng = model.engine(torch.device(DEV, 0))
    twin = eng.clone()
    assert twin.w is eng.w and twin.buf is not eng.buf
    res = {}
    for f in (1, 2, 3):
        stream = DepthMapStream(model, temperature=T, in_flight=f)
        assert len(stream.slots) == f + 1
        pending, got = [], []
        for s in items:
            pending.append(stream.submit(s.imgs, s.proj_matrices, s.depth_values))
            if len(pending) > stream.in_flight:
                got.append({k: v.clone() for k, v in stream.result(pending.pop(0)).items()})
        got += [{k: v.clone() for k, v in stream.result(t).items()} for t in pending]
        res[f] = got
    for f in (2, 3):
        assert len(res[f]) == len(items)
        for one, many in zip(res[1], res[f]):
            assert one.keys() == many.keys()
            for k in one:
                assert one[k].shape == many[k].shape
                err = O.rel_l1(many[k], one[k])
                assert err < _tol(k.split(".")[1]), (f, k, err)
    with pytest.raises(ValueError):
        DepthMapStream(model, depth=1, in_flight=2)


def test_forward_graph_matches_eager(pretrained_sd):
    """The CUDA-graph replay of the cascade returns what the launch-by-launch forward returns (several inputs through one
    captured graph; re-association noise of the atomically accumulated statistics only)."""
    cfg = dict(W=160, H=128, N=3, ndepths=(16, 8, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=192, interval=2.65)
    model = build(pretrained_sd, cfg["ndepths"], cfg["ratios"], torch.float16)
    eng = model.engine(torch.device(DEV, 0))
    for seed in range(3):
        s = synthetic.make_sample(cfg, "plane" if seed % 2 else "noise", seed=seed)
        args = (s.imgs.to(DEV), {k: v.to(DEV) for k, v in s.proj_matrices.items()}, s.depth_values.to(DEV), T)
        eager = {k: {kk: vv.clone() for kk, vv in v.items()} for k, v in eng.forward(*args).items() if isinstance(v, dict)}
        graphed = eng.forward_graph(*args)
        torch.cuda.synchronize()
        for st, maps in eager.items():
            for name, ref in maps.items():
                err = O.rel_l1(graphed[st][name].cpu(), ref.cpu())
                assert err < _tol(name), (seed, st, name, err)


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg5"])
def test_full_size_known_answer_and_storage_agreement(pretrained_sd, name):
    """BASELINE.json's full-size workloads (1600x1184 N=5 D=48/32/8; 1920x1056 N=7 D=64/32/8; 1600x1184 D=128/32/8), where the
    CPU oracle needs a minute per map: size-independent checks instead.  (1) Known answer: the photo-consistent slanted
    plane is recovered to about a millimetre at depths of 594-717 mm, with high confidence.  (2) The
    production path (fp16 storage, tcgen05 kernels) and the fp32-storage CUDA-core build -- two independent kernel sets,
    the second pinned to the live reference's goldens at the sizes the oracle can reach -- agree within north_star's
    1e-3 relative L1 on every stage's depth."""
    cfg = synthetic.CONFIGS[name]
    s = synthetic.make_sample(name, "plane", seed=0)
    outs = {}
    for storage in (torch.float16, torch.float32):
        model = build(pretrained_sd, cfg["ndepths"], cfg["ratios"], storage)
        out = run(model, s)
        outs[storage] = {k: {kk: vv.cpu() for kk, vv in v.items()} for k, v in out.items() if isinstance(v, dict)}
        del model, out
        torch.cuda.empty_cache()
    gt = s.gt_depth
    for st in range(len(cfg["ndepths"])):
        nm = f"stage{st + 1}"
        d16, d32 = outs[torch.float16][nm]["depth"], outs[torch.float32][nm]["depth"]
        rel = O.rel_l1(d16, d32)
        scale = gt.shape[-1] // d32.shape[-1]
        err32 = (d32 - gt[:, ::scale, ::scale]).abs().mean().item() if scale > 1 else (d32 - gt).abs().mean().item()
        err16 = (d16 - gt[:, ::scale, ::scale]).abs().mean().item() if scale > 1 else (d16 - gt).abs().mean().item()
        print(f"{name} {nm}: fp16-vs-fp32 storage depth rel-L1 {rel:.3e}; |depth - gt| fp32 {err32:.3f} mm, fp16 {err16:.3f} mm; "
              f"mean confidence {outs[torch.float16][nm]['photometric_confidence'].mean().item():.3f}")
        assert rel < DEPTH_REL_L1, (nm, rel)
    final16 = outs[torch.float16][f"stage{len(cfg['ndepths'])}"]
    # measured on B200: 0.24 / 0.15 / 0.18 mm and 0.97 / 0.97 / 0.96 (cfg2 / cfg3 / cfg5); the reference reaches 0.73 mm at 256x320
    assert (final16["depth"] - gt).abs().mean() < 0.5            # mm, of ~650 mm
    assert final16["photometric_confidence"].mean() > 0.9
