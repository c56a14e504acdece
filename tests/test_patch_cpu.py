"""Host logic of ``patch()`` against the LIVE reference package (oracle/_ref, imported from its archive): which names are
rebound at each level, that ``unpatch`` restores the reference, and that the reference's own constructor -- run on the patched
names -- yields a module tree with exactly the reference's state-dict keys and shapes (so its checkpoints load unchanged).
No kernel runs here; the executed drop-in is tests/test_gpu_parity.py::test_patch_drops_into_live_reference."""
import pytest
import torch

import cds_mvsnet_b200 as C
from oracle import ref_live

pytestmark = pytest.mark.skipif(not ref_live.available(), reason="oracle/_ref/reference_models.zip not shipped (run build())")


def test_levels_rebind_and_unpatch_restores():
    rmodel, rmodule, _, _ = ref_live.load()
    orig = {n: getattr(rmodel, n) for n in ("CDSMVSNet", "StageNet", "FeatureNet", "CostRegNet", "Refinement", "homo_warping_3D",
                                            "depth_regression", "conf_regression")}
    orig_dyn = rmodule.DynamicConv
    for level in C.PATCH_LEVELS:
        saved = C.patch(rmodel, rmodule, level=level)
        try:
            assert rmodel.homo_warping_3D is C.homo_warping_3D and rmodel.depth_regression is C.depth_regression
            assert rmodel.CostRegNet is C.CostRegNet and rmodule.DynamicConv is C.DynamicConv
            assert (rmodel.FeatureNet is C.FeatureNet) == (level != "leaf")
            assert (rmodel.StageNet is C.StageNet) == (level in ("stage", "model"))
            assert (rmodel.CDSMVSNet is C.CDSMVSNet) == (level == "model")
        finally:
            C.unpatch(saved)
        assert all(getattr(rmodel, n) is o for n, o in orig.items()) and rmodule.DynamicConv is orig_dyn
    with pytest.raises(ValueError):
        C.patch(rmodel, rmodule, level="everything")


@pytest.mark.parametrize("refine", [False, True])
@pytest.mark.parametrize("level", C.PATCH_LEVELS)
def test_reference_constructor_on_patched_names_keeps_state_dict(pretrained_sd, golden, level, refine):
    rmodel, rmodule, _, _ = ref_live.load()
    sd = dict(pretrained_sd)
    if refine:
        sd.update(golden("weights_refine_both_dtu_blended"))
    ref = ref_live.build_model(sd, (48, 32, 8), (4.0, 1.5, 0.75), refine=refine)
    want = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    saved = C.patch(rmodel, rmodule, level=level)
    try:
        m = ref_live.build_model(sd, (48, 32, 8), (4.0, 1.5, 0.75), refine=refine, rmodel=rmodel)   # strict load inside
    finally:
        C.unpatch(saved)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == want
    assert len(got) == (387 if refine else 356)
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]), k
    if level != "model":   # the reference's own driver class, holding CUDA-backed operators
        assert type(m).__module__ == "models.model"
        assert isinstance(m.cost_regularization[0], C.CostRegNet)


def test_cpu_tensor_is_refused_not_routed_to_a_fallback(pretrained_sd):
    m = C.CDSMVSNet(ndepths=(8,), depth_interals_ratio=(1.0,))
    m.eval()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 3, 64, 64), {"stage1": torch.zeros(1, 3, 2, 4, 4)}, torch.arange(8.0).unsqueeze(0))
