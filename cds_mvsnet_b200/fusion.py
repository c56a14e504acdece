"""Geometric-consistency filter on the CUDA kernels: drop-ins for the reference's ``fusion.py`` (the step after the
depth-inference path; driven by test.py:326-352).

    get_reproj(ref_depth, srcs_depth, ref_cam, srcs_cam) -> (reproj_xyd [n,v,3,h,w], in_range [n,v,1,h,w])   fusion.py:80-100
    vis_filter(ref_depth, reproj_xyd, in_range, img_dist_thresh, depth_thresh, vthresh) -> (masks, mask)      fusion.py:103-112
    ave_fusion(ref_depth, reproj_xyd, masks) -> ave [n,1,h,w]                                                  fusion.py:115-117
    prob_filter(ref_prob, prob_thresh) -> bool mask [n,1,h,w]                                                  fusion.py:69-77
    geometric_filter(...)   the whole of test.py:332-347 for one reference view in one pass (nothing but the outputs
                            touches HBM: the per-view (x, y, depth) maps are never materialised)

Same argument layout as the reference (cams are [.,2,4,4] = extrinsic, intrinsic in [1,:3,:3]); tensors must be CUDA
fp32.  No fallback: the arithmetic is in libcds_b200.so (csrc/fusion.cu).
"""
from __future__ import annotations

import ctypes

import torch

from ._lib import LIB, call, ptr


def _f32(t):
    if not t.is_cuda:
        raise RuntimeError("cds_mvsnet_b200.fusion needs CUDA tensors (there is no CPU fallback)")
    return t.to(torch.float32).contiguous()


def _mats(ref_cam, srcs_cam):
    n, v = srcs_cam.shape[:2]
    if tuple(ref_cam.shape) != (n, 2, 4, 4) or tuple(srcs_cam.shape) != (n, v, 2, 4, 4):
        raise AssertionError(f"cams must be ref [n,2,4,4] and srcs [n,v,2,4,4], got {tuple(ref_cam.shape)} / {tuple(srcs_cam.shape)}")
    rc, sc = _f32(ref_cam), _f32(srcs_cam)
    mats = torch.empty(LIB.load().cds_fusion_mats_floats(n, v), dtype=torch.float32, device=rc.device)
    call("cds_fusion_setup", ptr(rc), ptr(sc), n, v, ptr(mats))
    return mats


def get_reproj(ref_depth, srcs_depth, ref_cam, srcs_cam):
    n, v, _, h, w = srcs_depth.shape
    rd, sd = _f32(ref_depth), _f32(srcs_depth)
    mats = _mats(ref_cam, srcs_cam)
    xyd = torch.empty(n, v, 3, h, w, dtype=torch.float32, device=rd.device)
    inr = torch.empty(n, v, 1, h, w, dtype=torch.float32, device=rd.device)
    call("cds_geometric_filter", ptr(rd), ptr(sd), ptr(mats), n, v, h, w, 0.0, 0.0, 0.0, ptr(xyd), ptr(inr), None, None, None, None)
    return xyd, inr


def vis_filter(ref_depth, reproj_xyd, in_range, img_dist_thresh, depth_thresh, vthresh):
    n, v, _, h, w = reproj_xyd.shape
    rd, xyd, inr = _f32(ref_depth), _f32(reproj_xyd), _f32(in_range)
    masks = torch.empty(n, v, 1, h, w, dtype=torch.float32, device=rd.device)
    mask = torch.empty(n, 1, h, w, dtype=torch.uint8, device=rd.device)
    call("cds_vis_filter", ptr(rd), ptr(xyd), ptr(inr), None, n, v, h, w, float(img_dist_thresh), float(depth_thresh), float(vthresh),
         ptr(masks), ptr(mask), None)
    return masks, mask.bool()


def ave_fusion(ref_depth, reproj_xyd, masks):
    n, v, _, h, w = reproj_xyd.shape
    rd, xyd, mk = _f32(ref_depth), _f32(reproj_xyd), _f32(masks)
    ave = torch.empty(n, 1, h, w, dtype=torch.float32, device=rd.device)
    call("cds_vis_filter", ptr(rd), ptr(xyd), None, ptr(mk), n, v, h, w, 0.0, 0.0, 0.0, None, None, ptr(ave))
    return ave


def _thresholds(prob_thresh, C):
    th = [float(p) for p in prob_thresh]
    if not 1 <= len(th) <= min(C, 4):
        raise AssertionError(f"prob_thresh must have 1..{min(C, 4)} entries (got {len(th)})")
    return (ctypes.c_float * len(th))(*th), len(th)


def prob_filter(ref_prob, prob_thresh, greater=True):
    n, C, h, w = ref_prob.shape
    th, k = _thresholds(prob_thresh, C)
    pr = _f32(ref_prob[:, :k])
    mask = torch.empty(n, 1, h, w, dtype=torch.uint8, device=pr.device)
    call("cds_prob_filter", ptr(pr), th, n, k, h, w, ptr(mask), None)
    return mask.bool()


def geometric_filter(ref_depth, srcs_depth, ref_cam, srcs_cam, img_dist_thresh, depth_thresh, vthresh, ref_conf=None, srcs_conf=None,
                     prob_thresh=None, want_reproj=False):
    """test.py:332-347 for one batch of reference views: optional confidence masking of the source depths, reprojection,
    visibility filter, average fusion and back-projection, fused.  Returns dict(vis_mask, ave, points, masks[, prob_mask,
    final_mask, reproj_xyd, in_range])."""
    n, v, _, h, w = srcs_depth.shape
    rd = _f32(ref_depth)
    sd = _f32(srcs_depth)
    out = {}
    if prob_thresh is not None:
        if sd.data_ptr() == srcs_depth.data_ptr():
            sd = sd.clone()                       # the reference masks its own copy (test.py:335 multiplies in place on the batch)
        C = srcs_conf.shape[2]
        th, k = _thresholds(prob_thresh, C)
        sc = _f32(srcs_conf[:, :, :k]).reshape(n * v, k, h, w)
        call("cds_prob_filter", ptr(sc), th, n * v, k, h, w, None, ptr(sd))
        out["prob_mask"] = prob_filter(ref_conf, prob_thresh)
    mats = _mats(ref_cam, srcs_cam)
    dev, f32 = rd.device, torch.float32
    masks = torch.empty(n, v, 1, h, w, dtype=f32, device=dev)
    mask = torch.empty(n, 1, h, w, dtype=torch.uint8, device=dev)
    ave = torch.empty(n, 1, h, w, dtype=f32, device=dev)
    pts = torch.empty(n, 3, h, w, dtype=f32, device=dev)
    xyd = torch.empty(n, v, 3, h, w, dtype=f32, device=dev) if want_reproj else None
    inr = torch.empty(n, v, 1, h, w, dtype=f32, device=dev) if want_reproj else None
    call("cds_geometric_filter", ptr(rd), ptr(sd), ptr(mats), n, v, h, w, float(img_dist_thresh), float(depth_thresh), float(vthresh),
         ptr(xyd), ptr(inr), ptr(masks), ptr(mask), ptr(ave), ptr(pts))
    out.update(masks=masks, vis_mask=mask.bool(), ave=ave, points=pts)
    if want_reproj:
        out.update(reproj_xyd=xyd, in_range=inr)
    if "prob_mask" in out:
        out["final_mask"] = out["prob_mask"] & out["vis_mask"]
    return out
