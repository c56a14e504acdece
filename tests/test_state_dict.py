"""Drop-in modules keep the reference's state-dict keys/shapes; derived weight packing is exact (CPU)."""
import torch

import cds_mvsnet_b200 as C
from cds_mvsnet_b200 import synthetic, weights as W


def test_pretrained_keys_load_strict(pretrained_sd):
    m = C.CDSMVSNet(ndepths=(48, 32, 8), depth_interals_ratio=(4.0, 1.5, 0.75))
    res = m.load_state_dict(pretrained_sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert len(m.state_dict()) == len(pretrained_sd)


def test_random_state_dict_matches_module():
    m = C.CDSMVSNet()
    sd = synthetic.random_state_dict()
    ref = m.state_dict()
    assert set(sd) == set(ref)
    for k in sd:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k


def test_bn_fold_and_layouts(pretrained_sd):
    sd = pretrained_sd
    l = W.pack_conv3d(sd, "cost_regularization.0.conv1", False, "cpu")
    w = sd["cost_regularization.0.conv1.conv.weight"]
    scale = sd["cost_regularization.0.conv1.bn.weight"] / torch.sqrt(sd["cost_regularization.0.conv1.bn.running_var"] + 1e-5)
    # tap (kd,kh,kw)=(2,0,1), cin 3, cout 5
    assert torch.isclose(l.w[(2 * 3 + 0) * 3 + 1, 3, 5], w[5, 3, 2, 0, 1] * scale[5], rtol=1e-6)
    d = W.pack_conv3d(sd, "cost_regularization.1.conv9", True, "cpu")
    wt = sd["cost_regularization.1.conv9.conv.weight"]          # [Cin, Cout, ...]
    sc = sd["cost_regularization.1.conv9.bn.weight"] / torch.sqrt(sd["cost_regularization.1.conv9.bn.running_var"] + 1e-5)
    assert d.cin == 32 and d.cout == 16
    assert torch.isclose(d.w[(1 * 3 + 2) * 3 + 0, 7, 9], wt[7, 9, 1, 2, 0] * sc[9], rtol=1e-6)
    dyn = W.pack_dynamic_conv(sd, "feature.conv00.conv", 3, 8, (3, 7, 11), "cpu")
    assert dyn.w_att.shape == (179, 3, 4) and dyn.w_conv.shape == (179, 3, 8) and dyn.bias is None
    # second branch (7x7) starts after the 9 taps of the 3x3 branch
    a = sd["feature.conv00.conv.att_convs.1.weight"]
    assert torch.equal(dyn.w_att[9 + 2 * 7 + 4, 1, :3], a[:, 1, 2, 4])
    assert W.pack_dynamic_conv(sd, "feature.out1", 32, 32, (1, 3), "cpu").bias.shape == (2, 32)


def test_operand_images_match_the_library_layouts(pretrained_sd):
    """The host packers and the kernels agree on the size of every tensor-core operand image (asked of the library, no GPU),
    and the folded layouts put a known weight where the kernels' index arithmetic expects it."""
    from cds_mvsnet_b200 import _lib
    lib = _lib.LIB.load()
    sd = pretrained_sd
    # strided 3x3 convs on the row-streaming kernel: [kernel row][image][k-chunk][2*Cout][8 k]
    for name, ci, co in (("downsample1", 8, 16), ("downsample2", 16, 32)):
        w = sd[f"feature.{name}.conv.weight"]                                   # [Cout, Cin, 3, 3]
        img = W.pack_conv2d_s2rows(w.permute(2, 3, 1, 0).reshape(9, ci, co).contiguous())
        assert img.numel() == lib.cds_conv2d_3x3s2_rows_weight_halfs(ci, co)
        hi = lambda t: t.half().float()
        if ci == 8:   # image 0 = taps (dx 0, dx 2), image 1 = (dx 1, dx 1)
            assert img[2, 0, 1, 5 // 8, 5 % 8, 3] == hi(w[5, 3, 2, 2]).half()
            assert img[1, 1, 0, 9 // 8, 9 % 8, 6] == img[1, 1, 1, 9 // 8, 9 % 8, 6] == hi(w[9, 6, 1, 1]).half()
            # columns [Cout, 2 Cout): what fp16 rounding of the weight dropped
            assert img[0, 0, 0, (co + 2) // 8, (co + 2) % 8, 1] == (w[2, 1, 0, 0] - hi(w[2, 1, 0, 0])).half()
        else:         # image dx, k-chunk = channel chunk
            assert img[0, 2, 1, 20 // 8, 20 % 8, 4] == hi(w[20, 12, 0, 2]).half()
    # visibility net: kernel rows folded into N: column group g <-> kernel row 2 - g
    wgt, fp = W.pack_visnet_tc(sd, "stage_net.vis.2", "cpu")
    assert wgt.numel() == lib.cds_visnet_tc_weight_halfs() and fp.numel() == 65
    scale = sd["stage_net.vis.2.1.bn.weight"] / torch.sqrt(sd["stage_net.vis.2.1.bn.running_var"] + 1e-5)
    l2 = wgt[2 * 2 * 48 * 8:].reshape(-1)[:3 * 2 * 48 * 8].reshape(3, 2, 48, 8)   # layer 2: [kw][k-chunk][column][k]
    want = (sd["stage_net.vis.2.1.conv.weight"][7, 11, 0, 2].double() * scale[7].double()).half()   # cout 7, cin 11, kh 0, kw 2
    assert l2[2, 1, 2 * 16 + 7, 3] == want                                                         # group 2 <-> kernel row 0
    # row-folded DynamicConv images
    import ctypes
    for name in ("conv00", "conv01", "conv10", "conv20", "out2", "out3"):
        ci, co, ks, pre = W.DYN_LAYERS[name]
        dw = W.pack_dynamic_conv(sd, pre, ci, co, ks, "cpu")
        kz = (ctypes.c_int * len(ks))(*ks)
        assert W.pack_dynamic_conv_kh(dw).numel() == lib.cds_dynamic_conv_kh_weight_halfs(max(8, ci), co, len(ks), kz)


def test_training_mode_is_refused():
    import pytest
    m = C.CDSMVSNet()
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 3, 3, 64, 64), {}, torch.zeros(1, 4))
    with pytest.raises(NotImplementedError):
        C.CDSMVSNet(share_cr=True)


def test_refine_state_dict_keys(golden, pretrained_sd):
    """refine=True adds the reference's refine_network.* entries (models/module.py:318-335): a full checkpoint loads strictly."""
    rsd = golden("weights_refine_both_dtu_blended")
    m = C.CDSMVSNet(refine=True, ndepths=(48, 32, 8), depth_interals_ratio=(4.0, 1.5, 0.75))
    sd = dict(pretrained_sd)
    sd.update(rsd)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert len(rsd) == 31 and all(k.startswith("refine_network.") for k in rsd)
    folded = W.pack_refinement(sd, "refine_network", "cpu")
    w, b = folded["deconv"]
    scale = sd["refine_network.bn.weight"] / torch.sqrt(sd["refine_network.bn.running_var"] + 1e-5)
    assert torch.isclose(w[3, 5, 1, 2], sd["refine_network.deconv.weight"][3, 5, 1, 2] * scale[5], rtol=1e-6)
    assert folded["res"].shape == (8, 9) and folded["conv3"][0].shape == (8, 16, 3, 3)
