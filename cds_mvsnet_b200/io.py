"""On-disk outputs of the depth-inference job, format-compatible with the reference: PFM maps
(datasets/data_io.py:6-71), `*_cam.txt` (test.py:132-149) and the per-view layout written at test.py:218-248
(`depth_est/*.pfm`, 3-channel `confidence/*.pfm` = the three stages' confidences nearest-resized to the depth map, `cams/*_cam.txt`).
Host-side numpy only; the files produced are byte-identical to the reference's (tests/test_io.py, fixture written by the
reference's own save_pfm / write_cam)."""
from __future__ import annotations

import os
import re
import sys

import numpy as np


def read_pfm(filename):
    """-> (array [H,W] or [H,W,3] float32 top row first, scale)   (datasets/data_io.py:6-41)"""
    with open(filename, "rb") as f:
        header = f.readline().decode("utf-8").rstrip()
        if header not in ("PF", "Pf"):
            raise Exception("Not a PFM file.")
        color = header == "PF"
        m = re.match(r"^(\d+)\s(\d+)\s$", f.readline().decode("utf-8"))
        if not m:
            raise Exception("Malformed PFM header.")
        width, height = map(int, m.groups())
        scale = float(f.readline().rstrip())
        endian = "<" if scale < 0 else ">"
        data = np.fromfile(f, endian + "f")
    shape = (height, width, 3) if color else (height, width)
    return np.flipud(np.reshape(data, shape)), abs(scale)


def save_pfm(filename, image, scale=1):
    """image float32 [H,W], [H,W,1] or [H,W,3]   (datasets/data_io.py:44-71)"""
    image = np.flipud(image)
    if image.dtype.name != "float32":
        raise Exception("Image dtype must be float32.")
    if image.ndim == 3 and image.shape[2] == 3:
        color = True
    elif image.ndim == 2 or (image.ndim == 3 and image.shape[2] == 1):
        color = False
    else:
        raise Exception("Image must have H x W x 3, H x W x 1 or H x W dimensions.")
    endian = image.dtype.byteorder
    if endian == "<" or (endian == "=" and sys.byteorder == "little"):
        scale = -scale
    with open(filename, "wb") as f:
        f.write(b"PF\n" if color else b"Pf\n")
        f.write("{} {}\n".format(image.shape[1], image.shape[0]).encode("utf-8"))
        f.write(("%f\n" % scale).encode("utf-8"))
        image.tofile(f)


def write_cam(filename, cam):
    """cam [2,4,4] = (extrinsic, intrinsic; row [1,3] = depth_min, interval, ndepth, depth_max).  Same bytes as the text file
    of test.py:132-149: every number as ``str()`` of the array element, a trailing blank after each matrix entry."""
    cam = np.asarray(cam)
    row = lambda r: "".join(str(v) + " " for v in r)
    lines = ["extrinsic"] + [row(cam[0, i]) for i in range(4)] + ["", "intrinsic"] + [row(cam[1, i, :3]) for i in range(3)]
    lines += ["", " ".join(str(v) for v in cam[1, 3])]
    with open(filename, "w") as f:
        f.write("\n".join(lines) + "\n")


def resize_nearest(a, h, w):
    """cv2.resize(a, (w, h), interpolation=cv2.INTER_NEAREST) for a [H0,W0(,C)] array: source index = floor(dst * src / dst_size)."""
    h0, w0 = a.shape[:2]
    ys = np.minimum(np.floor(np.arange(h) * (h0 / h)).astype(np.int64), h0 - 1)
    xs = np.minimum(np.floor(np.arange(w) * (w0 / w)).astype(np.int64), w0 - 1)
    return a[ys][:, xs]


def save_view(outdir, filename, depth_est, stage_confidences, cam):
    """One reference view as test.py:218-248 writes it.  filename: the dataset's pattern, e.g. 'scan1/{}/00000000{}';
    depth_est [H,W] (the refined depth); stage_confidences: the three stages' [h_s,w_s] maps; cam [2,4,4]."""
    paths = {k: os.path.join(outdir, filename.format(k, ext)) for k, ext in (("depth_est", ".pfm"), ("confidence", ".pfm"), ("cams", "_cam.txt"))}
    for p in paths.values():
        os.makedirs(os.path.dirname(p), exist_ok=True)
    depth_est = np.ascontiguousarray(depth_est, dtype=np.float32)
    save_pfm(paths["depth_est"], depth_est)
    h, w = depth_est.shape
    conf = np.stack([resize_nearest(np.asarray(c, dtype=np.float32), h, w) for c in stage_confidences]).transpose([1, 2, 0])
    save_pfm(paths["confidence"], np.ascontiguousarray(conf))
    write_cam(paths["cams"], cam)
    return paths
