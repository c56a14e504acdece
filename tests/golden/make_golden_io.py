"""Files written by the LIVE reference's save_pfm (datasets/data_io.py) and write_cam (test.py:132-149) for small arrays, kept as
byte strings: the fixture of tests/test_io.py.    python tests/golden/make_golden_io.py -> tests/golden/io_files.npz"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CDS_REF_PATH", "/root/reference")
sys.path.insert(0, REF)
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_data_io", os.path.join(REF, "datasets", "data_io.py"))
src = open(os.path.join(REF, "datasets", "data_io.py")).read().split("import random, cv2")[0]      # the PFM part needs no cv2
ns = {}
exec(compile(src, "data_io_pfm", "exec"), ns)

# write_cam lives in test.py, which imports plyfile etc.: take the function's source only
tsrc = open(os.path.join(REF, "test.py")).read()
start = tsrc.index("def write_cam(file, cam):")
end = tsrc.index("\ndef ", start + 10)
exec(compile(tsrc[start:end], "test_write_cam", "exec"), ns)

rng = np.random.default_rng(0)
depth = (425 + 500 * rng.random((5, 7))).astype(np.float32)
conf = rng.random((5, 7, 3)).astype(np.float32)
cam = np.zeros((2, 4, 4), dtype=np.float32)
cam[0] = np.eye(4) + 0.01 * rng.standard_normal((4, 4))
cam[1, :3, :3] = [[1920.5, 0, 800.25], [0, 1920.5, 592.125], [0, 0, 1]]
cam[1, 3] = [425.0, 2.65, 192, 931.15]
with tempfile.TemporaryDirectory() as d:
    ns["save_pfm"](os.path.join(d, "a.pfm"), depth)
    ns["save_pfm"](os.path.join(d, "b.pfm"), conf)
    ns["write_cam"](os.path.join(d, "c.txt"), cam)
    files = {k: np.frombuffer(open(os.path.join(d, n), "rb").read(), dtype=np.uint8) for k, n in (("depth_pfm", "a.pfm"), ("conf_pfm", "b.pfm"), ("cam_txt", "c.txt"))}
    back, scale = ns["read_pfm"](os.path.join(d, "a.pfm"))
    assert np.array_equal(back, depth) and scale == 1.0
np.savez_compressed(os.path.join(HERE, "io_files.npz"), depth=depth, conf=conf, cam=cam, **files)
print({k: v.size for k, v in files.items()})
