"""CPU restatement of the reference's geometric-consistency filter (the step right after the depth-inference path:
fusion.py:49-114, driven by test.py:326-352).  TEST INFRASTRUCTURE ONLY: imported by tests/ and the golden generator, never by
the product path.  Pinned against outputs of the live reference (tests/golden/make_golden_fusion.py -> fusion_*.npz).

Conventions of the reference kept verbatim: pixel centres at +0.5 (fusion.py:7-12); cam = [extrinsic 4x4, intrinsic in
[1,:3,:3]] (fusion.py:22,30); every homogeneous division adds 1e-9 (fusion.py:23,32,38,44,46); the warp coordinate is
divided by width/height, mapped to [-1,1], clamped to +-1.1 and sampled with grid_sample(bilinear, zeros,
align_corners=True) (fusion.py:57-64) -- i.e. at pixel position coord * (W-1)/W, not at coord - 0.5."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def pixel_grids(h, w):   # fusion.py:7-12  -> [h, w, 3, 1]
    x = (torch.arange(w, dtype=torch.float32) + 0.5).repeat(h, 1)
    y = (torch.arange(h, dtype=torch.float32) + 0.5).repeat(w, 1).t()
    return torch.stack([x, y, torch.ones_like(x)], dim=-1).unsqueeze(-1)


def img2cam(idx_img, depth, cam):   # fusion.py:22-27 ; nhw31, n1hw -> nhw41
    c = cam[:, 1:2, :3, :3].unsqueeze(1).inverse() @ idx_img
    c = c / (c[..., -1:, :] + 1e-9) * depth.permute(0, 2, 3, 1).unsqueeze(4)
    return torch.cat([c, torch.ones_like(c[..., -1:, :])], dim=-2)


def cam2world(c, cam):   # fusion.py:30-33
    wv = cam[:, 0:1, ...].unsqueeze(1).inverse() @ c
    return wv / (wv[..., -1:, :] + 1e-9)


def world2cam(wv, cam):   # fusion.py:36-39
    c = cam[:, 0:1, ...].unsqueeze(1) @ wv
    return c / (c[..., -1:, :] + 1e-9)


def cam2img(c, cam):   # fusion.py:42-47
    c3 = c[..., :3, :] / (c[..., 3:4, :] + 1e-9)
    i = cam[:, 1:2, :3, :3].unsqueeze(1) @ c3
    return i / (i[..., -1:, :] + 1e-9)


def project_img(src_img, dst_depth, src_cam, dst_cam):   # fusion.py:49-66
    h, w = src_img.shape[-2:]
    g = pixel_grids(h, w).unsqueeze(0)
    img = cam2img(world2cam(cam2world(img2cam(g, dst_depth, dst_cam), dst_cam), src_cam), src_cam)
    coord = img[..., :2, 0].clone()
    coord[..., 0] /= w
    coord[..., 1] /= h
    coord = (coord * 2 - 1).clamp(-1.1, 1.1)
    in_range = ((-1 <= coord[..., 0]) & (coord[..., 0] <= 1) & (-1 <= coord[..., 1]) & (coord[..., 1] <= 1)).to(src_img.dtype).unsqueeze(1)
    return F.grid_sample(src_img, coord, mode="bilinear", padding_mode="zeros", align_corners=True), in_range


def prob_filter(ref_prob, prob_thresh):   # fusion.py:69-77 ; n c h w, thresholds per channel
    mask = None
    for i, p in enumerate(prob_thresh):
        m = ref_prob[:, [i]] > p
        mask = m if mask is None else (mask & m)
    return mask


def get_reproj(ref_depth, srcs_depth, ref_cam, srcs_cam):   # fusion.py:80-100
    n, v, _, h, w = srcs_depth.shape
    sd = srcs_depth.reshape(n * v, 1, h, w)
    sc = srcs_cam.reshape(n * v, 2, 4, 4)
    rd = ref_depth.unsqueeze(1).repeat(1, v, 1, 1, 1).reshape(n * v, 1, h, w)
    rc = ref_cam.unsqueeze(1).repeat(1, v, 1, 1, 1).reshape(n * v, 2, 4, 4)
    g = pixel_grids(h, w).unsqueeze(0)
    s_cam = img2cam(g, sd, sc)
    s2r_cam = world2cam(cam2world(s_cam, sc), rc)
    s2r_img = cam2img(s2r_cam, rc)
    xyd = torch.cat([s2r_img[..., :2, 0], s2r_cam[..., 2:3, 0]], dim=-1).permute(0, 3, 1, 2)
    reproj, in_range = project_img(xyd, rd, sc, rc)
    return reproj.reshape(n, v, 3, h, w), in_range.reshape(n, v, 1, h, w)


def vis_filter(ref_depth, reproj_xyd, in_range, img_dist_thresh, depth_thresh, vthresh):   # fusion.py:103-112
    n, v, _, h, w = reproj_xyd.shape
    xy = pixel_grids(h, w).permute(3, 2, 0, 1).unsqueeze(1)[:, :, :2]
    dist = (reproj_xyd[:, :, :2] - xy).norm(dim=2, keepdim=True) < img_dist_thresh
    rd = ref_depth.unsqueeze(1)
    dep = (rd - reproj_xyd[:, :, 2:]).abs() < (torch.max(rd, reproj_xyd[:, :, 2:]) * depth_thresh)
    masks = torch.min(torch.min(in_range, dist.to(ref_depth.dtype)), dep.to(ref_depth.dtype))
    return masks, masks.sum(dim=1) >= (vthresh - 1.1)


def ave_fusion(ref_depth, reproj_xyd, masks):   # fusion.py:115-117
    return ((reproj_xyd[:, :, 2:] * masks).sum(dim=1) + ref_depth) / (masks.sum(dim=1) + 1)


def back_project(depth, cam):   # test.py:345-347 -> world points [n, 3, h, w]
    h, w = depth.shape[-2:]
    g = pixel_grids(h, w).unsqueeze(0)
    return cam2world(img2cam(g, depth, cam), cam)[..., :3, 0].permute(0, 3, 1, 2)
