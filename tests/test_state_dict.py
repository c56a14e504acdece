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
