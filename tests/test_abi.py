"""The C-ABI library loads and exports every symbol include/cds_b200.h declares (CPU box: no compute calls)."""
import ctypes
import os

import pytest

from cds_mvsnet_b200 import _lib


def test_header_parses():
    protos = _lib.parse_header()
    assert len(protos) >= 20
    for must in ("cds_homo_warp", "cds_costvol_entropy", "cds_costvol_aggregate", "cds_visnet", "cds_conv3d_k3",
                 "cds_deconv3d_k3s2", "cds_softmax_regress", "cds_dynamic_conv", "cds_instnorm_act",
                 "cds_depth_hypotheses", "cds_version", "cds_last_error_string"):
        assert must in protos, must


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.fail(f"{_lib.LIB_PATH} missing: run `python -m cds_mvsnet_b200.build`")
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for name in _lib.parse_header():
        assert hasattr(dll, name), f"{name} declared in include/cds_b200.h but not exported"
    dll.cds_version.restype = ctypes.c_int
    assert dll.cds_version() >= 100
    dll.cds_visnet_weight_floats.restype = ctypes.c_int
    assert dll.cds_visnet_weight_floats() == 4961


def test_argument_checks_do_not_need_a_gpu():
    dll = _lib.LIB.load()
    # null pointers are rejected before any CUDA call
    assert dll.cds_homo_warp(None, None, None, 0, 1, 8, 4, 8, 8, None, None) == -1
    assert b"null pointer" in dll.cds_last_error_string()
    assert dll.cds_conv3d_k3(None, None, None, 1, 8, 8, 8, 8, 8, 1, 1, 1, None, None) == -1
    # training-slice entries: null pointers, bad shapes, the C % 4 rule of the workspace form
    assert dll.cds_homo_warp_backward(None, None, None, 0, 1, 8, 4, 8, 8, None, None, None) == -1
    assert dll.cds_homo_warp_backward(1, 1, 1, 0, 1, 8, 4, 1, 8, 1, None, None) == -2
    assert dll.cds_homo_warp_backward(1, 1, 1, 0, 1, 5, 4, 8, 8, 1, 1, None) == -3
    assert b"C % 4" in dll.cds_last_error_string()
    assert dll.cds_depth_regress_backward(1, None, None, 0, 1, 4, 8, 8, None, None, None) == -1
    assert dll.cds_stage_loss_forward(None, None, None, None, None, 1, 8, 8, None, None) == -1
    assert dll.cds_stage_loss_backward(1, 1, 1, 1, 1, None, None, 1, 8, 8, 1, None, None) == -1
    assert dll.cds_feat_loss_forward(1, 1, 1, 0, 4, 8, 8, 1, None) == -2


def test_training_ops_refuse_cpu_tensors():
    import torch

    import cds_mvsnet_b200 as C
    from cds_mvsnet_b200 import losses
    x = torch.zeros(1, 8, 4, 4, requires_grad=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        C.homo_warping_3D(x, torch.eye(4)[None], torch.eye(4)[None], torch.ones(1, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        C.depth_regression(torch.zeros(1, 2, 4, 4, requires_grad=True), torch.ones(1, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        losses.final_loss({f"stage{i}": {"depth": torch.zeros(1, 4, 4), "norm_curv": torch.zeros(1, 1, 4, 4)} for i in (1, 2, 3)},
                          {f"stage{i}": torch.zeros(1, 4, 4) for i in (1, 2, 3)}, {f"stage{i}": torch.ones(1, 4, 4) for i in (1, 2, 3)},
                          depth_interval=torch.ones(1))


def test_product_path_has_no_cpu_fallback():
    import torch

    import cds_mvsnet_b200 as C
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        C.homo_warping_3D(torch.zeros(1, 8, 4, 4), torch.eye(4)[None], torch.eye(4)[None], torch.ones(1, 2))


def test_product_does_not_import_oracle():
    import pathlib
    pkg = pathlib.Path(_lib.PKG)
    for f in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*")):
        txt = f.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt, f
