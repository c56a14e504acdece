"""PFM / cam writers are byte-identical to the reference's (fixture written by the live reference: make_golden_io.py)."""
import os

import numpy as np

from cds_mvsnet_b200 import io as cio


def test_files_are_byte_identical_to_the_reference(golden, tmp_path):
    g = {k: v.numpy() for k, v in golden("io_files").items()}
    cio.save_pfm(tmp_path / "a.pfm", g["depth"])
    cio.save_pfm(tmp_path / "b.pfm", g["conf"])
    cio.write_cam(tmp_path / "c.txt", g["cam"])
    for name, key in (("a.pfm", "depth_pfm"), ("b.pfm", "conf_pfm"), ("c.txt", "cam_txt")):
        assert open(tmp_path / name, "rb").read() == g[key].tobytes(), name
    d, s = cio.read_pfm(tmp_path / "a.pfm")
    assert np.array_equal(d, g["depth"]) and s == 1.0
    c, _ = cio.read_pfm(tmp_path / "b.pfm")
    assert np.array_equal(c, g["conf"])


def test_save_view_layout_and_nearest_resize(tmp_path):
    rng = np.random.default_rng(1)
    depth = rng.random((8, 12)).astype(np.float32)
    confs = [rng.random((2, 3)).astype(np.float32), rng.random((4, 6)).astype(np.float32), rng.random((8, 12)).astype(np.float32)]
    cam = np.zeros((2, 4, 4), dtype=np.float32)
    paths = cio.save_view(str(tmp_path), "scan9/{}/00000003{}", depth, confs, cam)
    assert paths["depth_est"].endswith(os.path.join("scan9", "depth_est", "00000003.pfm"))
    assert paths["cams"].endswith(os.path.join("scan9", "cams", "00000003_cam.txt"))
    conf, _ = cio.read_pfm(paths["confidence"])
    assert conf.shape == (8, 12, 3)
    assert np.array_equal(conf[:, :, 2], confs[2])
    assert np.array_equal(conf[:, :, 0], np.repeat(np.repeat(confs[0], 4, axis=0), 4, axis=1))   # exact 4x: floor(i / 4)
    assert np.array_equal(cio.resize_nearest(confs[1], 8, 12), np.repeat(np.repeat(confs[1], 2, axis=0), 2, axis=1))
