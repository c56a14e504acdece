"""CPU diagnostic (no GPU): which fp16 rounding points of the CUDA path cost how much depth error on the chaotic "noise"
input.  The oracle's feature extractor / stage net are re-run with fp16 rounding emulated at selectable points:

  store:<layers>   the raw (pre-norm) output of these feature layers is rounded to fp16 (activation storage)
  act:<layers>     the normalised + activated INPUT of these layers is rounded to fp16 (tensor-core A operand)
  w:<layers>       the feature-conv weights of these layers are rounded to fp16 (tensor-core B operand)
  fea              the three stage features (tanh outputs) are rounded to fp16
  vol              the aggregated cost volume is rounded to fp16
  reg              every regulariser activation is rounded to fp16
  s1src16 | s1srci16 | s1ref16 | s1refi16   the stage-1 source / reference features as the plane sweep reads them: fp16 or
                   16-bit fixed point (tanh output in [-1, 1], step 2^-15)

usage: python scripts/ablate_cpu.py [--seeds=0-5] name=spec[,spec...] ...   e.g.  all=store:*,act:*,w:*,fea,vol,reg
"""
import sys
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, ".")
from cds_mvsnet_b200 import synthetic
from oracle import oracle as O

torch.set_grad_enabled(False)
O.FAST_GATHER = True
z = np.load("tests/golden/weights_both_dtu_blended.npz"); SD = {k: torch.from_numpy(z[k]) for k in z.files}
LAYERS = ["conv00", "conv01", "downsample1", "conv10", "conv11", "downsample2", "conv20", "conv21", "out1", "inner1", "out2", "inner2", "out3"]
h16 = lambda t: t.to(torch.float16).to(torch.float32)


class Spec:
    def __init__(self, text):
        self.store, self.act, self.w, self.flags, self.reg1 = set(), set(), set(), set(), set()
        for item in [t for t in text.split(",") if t]:
            if ":" in item:
                kind, _, names = item.partition(":")
                names = LAYERS if names == "*" else names.split("+")
                getattr(self, kind).update(names)
            else:
                self.flags.add(item)


def feature_net(img, sd, epipole, T, sp):
    def wq(name, sd_):   # fp16-rounded feature weights of a layer (curvature weights carry their residual on the GPU: exact)
        if name not in sp.w:
            return sd_
        out = dict(sd_)
        for k, v in sd_.items():
            if (k.startswith(f"feature.{name}.") and ".convs." in k and k.endswith("weight")) or k == f"feature.{name}.conv.weight":
                out[k] = h16(v)
        return out

    def a_in(name, x):
        return h16(x) if name in sp.act else x

    def st(name, y):
        return h16(y) if name in sp.store else y

    def dyn(name, x, e, prefix=None, norm=True):
        y, nc = O.dynamic_conv(a_in(name, x), wq(name, sd), prefix or f"feature.{name}.conv", O.FEATURE_DYN_KSIZES[name], e, T)
        return st(name, y), nc

    def act(y):
        return F.leaky_relu(O.instance_norm(y), 0.1)

    def plain(name, x, stride=1, padding=0):
        return st(name, F.conv2d(a_in(name, x), wq(name, sd)[f"feature.{name}.conv.weight"], stride=stride, padding=padding))

    e1, e2 = epipole / 2, epipole / 4
    # image: the GPU feeds hi + lo planes (exact), so conv00's input is never rounded
    r00, n00 = dyn("conv00", img, epipole); c00 = act(r00)
    r01, n01 = dyn("conv01", c00, epipole); c01 = act(r01)
    d1 = act(plain("downsample1", c01, 2, 1))
    r10, n10 = dyn("conv10", d1, e1); c10 = act(r10)
    r11, n11 = dyn("conv11", c10, e1); c11 = act(r11)
    d2 = act(plain("downsample2", c11, 2, 1))
    r20, n20 = dyn("conv20", d2, e2); c20 = act(r20)
    r21, n21 = dyn("conv21", c20, e2); c21 = act(r21)
    fq = (lambda t: h16(t)) if ("fea" in sp.flags or "fea23" in sp.flags) else (lambda t: t)
    fq1 = (lambda t: h16(t)) if "fea" in sp.flags else (lambda t: t)
    out = {}
    o1, n22 = dyn("out1", c21, e2, "feature.out1")
    out["stage1"] = (fq1(torch.tanh(O.instance_norm(o1))), (n20 ** 2 + n21 ** 2 + n22 ** 2) / 3, n22.abs())
    i1 = act(plain("inner1", torch.cat((O._up2_nearest(c21), c11), 1)))
    o2, n12 = dyn("out2", i1, e1, "feature.out2")
    o2 = fq(torch.tanh(O.instance_norm(o2)))
    out["stage2"] = (o2, (n10 ** 2 + n11 ** 2 + n12 ** 2) / 3, n12.abs())
    i2 = act(plain("inner2", torch.cat((O._up2_nearest(o2), c01), 1)))
    o3, n02 = dyn("out3", i2, epipole, "feature.out3")
    out["stage3"] = (fq(torch.tanh(O.instance_norm(o3))), (n00 ** 2 + n01 ** 2 + n02 ** 2) / 3, n02.abs())
    return out


def forward(s, cfg, sp):
    orig = (O.feature_net, O._cbr3d, O._dbr3d, O.cost_reg_net)
    O.feature_net = lambda img, sd, e, T: feature_net(img, sd, e, T, sp)
    s1 = lambda p: p.startswith("cost_regularization.0")
    if "reg" in sp.flags or "reg23" in sp.flags:
        skip1 = "reg" not in sp.flags
        O._cbr3d = lambda x, sd, p, stride=1: (orig[1](x, sd, p, stride) if skip1 and s1(p) else h16(orig[1](x, sd, p, stride)))
        O._dbr3d = lambda x, sd, p: (orig[2](x, sd, p) if skip1 and s1(p) else h16(orig[2](x, sd, p)))
    if sp.reg1:   # fp16 rounding of the OUTPUT of the named stage-1 regulariser layers only
        hit = lambda p: s1(p) and p.rsplit(".", 1)[1] in sp.reg1
        O._cbr3d = lambda x, sd, p, stride=1: (h16(orig[1](x, sd, p, stride)) if hit(p) else orig[1](x, sd, p, stride))
        O._dbr3d = lambda x, sd, p: (h16(orig[2](x, sd, p)) if hit(p) else orig[2](x, sd, p))
    if "vol" in sp.flags or "vol23" in sp.flags:
        skipv = "vol" not in sp.flags
        O.cost_reg_net = lambda x, sd, prefix: orig[3](x if skipv and s1(prefix) else h16(x), sd, prefix)
    orig_stage = O.stage_net
    q16 = {"16": h16, "i16": lambda t: torch.round(t * 32768.0).clamp(-32768, 32767) / 32768.0}
    s1 = {side: q16[f[len("s1" + side):]] for f in sp.flags for side in ("src", "ref") if f.startswith("s1" + side)}
    if s1:   # stage-1 features as the aggregate sweep sees them: fp16 ("s1src16") or 16-bit fixed point ("s1srci16") per side
        def stage_net(features, proj, samples, sd, stage_idx, *a, **k):
            if stage_idx == 0:
                features = [{side: ((s1[side](f[side][0]) if side in s1 else f[side][0]),) + tuple(f[side][1:]) for side in ("ref", "src")}
                            for f in features]
            return orig_stage(features, proj, samples, sd, stage_idx, *a, **k)
        O.stage_net = stage_net
    try:
        return O.cdsmvsnet_forward(SD, s.imgs, s.proj_matrices, s.depth_values, cfg["ndepths"], cfg["ratios"], 0.01)
    finally:
        O.feature_net, O._cbr3d, O._dbr3d, O.cost_reg_net = orig
        O.stage_net = orig_stage


if __name__ == "__main__":
    seeds, variants, hw, nv = list(range(6)), {}, (128, 160), 4
    for a in sys.argv[1:]:
        if a.startswith("--seeds="):
            lo, hi = a[8:].split("-"); seeds = list(range(int(lo), int(hi) + 1))
        elif a.startswith("--hw="):
            hw = tuple(int(v) for v in a[5:].split("x"))
        elif a.startswith("--n="):
            nv = int(a[4:])
        else:
            name, _, text = a.partition("="); variants[name] = Spec(text)
    cfg = dict(W=hw[1], H=hw[0], N=nv, ndepths=(48, 32, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=192, interval=2.65)
    acc = {k: [] for k in variants}
    for seed in seeds:
        s = synthetic.make_sample(cfg, "noise", seed=seed)
        ref = forward(s, cfg, Spec(""))
        for name, sp in variants.items():
            out = forward(s, cfg, sp)
            acc[name].append([O.rel_l1(out[f"stage{i}"]["depth"], ref[f"stage{i}"]["depth"]) for i in (1, 2, 3)])
        print(seed, {k: "%.2e" % acc[k][-1][2] for k in acc}, flush=True)
    for name, rows in acc.items():
        r = np.array(rows)
        print(f"{name:28s} stage3 mean {r[:, 2].mean():.2e} max {r[:, 2].max():.2e} | per-stage mean {np.array2string(r.mean(0), precision=6)}")
