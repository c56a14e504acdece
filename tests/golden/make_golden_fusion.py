"""Golden outputs of the LIVE reference's geometric-consistency filter (fusion.py) on the synthetic fusion sample.
Run in the build container (the reference tree does not travel to the GPU box):

    python tests/golden/make_golden_fusion.py        # -> tests/golden/fusion_small.npz

The reference's get_pixel_grids calls .cuda(); there is no GPU here, so Tensor.cuda is patched to the identity for the
duration of the run (the arithmetic is the same ATen code on the CPU)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("CDS_REF_PATH", "/root/reference"))

from cds_mvsnet_b200 import synthetic  # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self
import fusion  # noqa: E402  (the reference's own module)

PROB_T, DISP_T, DEPTH_T, VIEW_T = (0.1, 0.1, 0.1), 1.0, 0.01, 3


def run(sample):
    s = {k: v.clone() for k, v in sample.items()}
    for i in range(s["src_depths"].size(1)):                                   # test.py:333-335
        s["src_depths"][:, i] *= fusion.prob_filter(s["src_confs"][:, i], PROB_T).float()
    prob_mask = fusion.prob_filter(s["ref_conf"], PROB_T)                      # test.py:337
    xyd, in_range = fusion.get_reproj(s["ref_depth"], s["src_depths"], s["ref_cam"], s["src_cams"])
    masks, mask = fusion.vis_filter(s["ref_depth"], xyd, in_range, DISP_T, DEPTH_T, VIEW_T)
    ave = fusion.ave_fusion(s["ref_depth"], xyd, masks)
    final = fusion.bin_op_reduce([prob_mask, mask], torch.min)
    g = fusion.get_pixel_grids(*ave.size()[-2:]).unsqueeze(0)
    pts = fusion.idx_cam2world(fusion.idx_img2cam(g, ave, s["ref_cam"]), s["ref_cam"])[..., :3, 0].permute(0, 3, 1, 2)
    return dict(reproj_xyd=xyd, in_range=in_range, masks=masks, vis_mask=mask, ave=ave, final_mask=final, points=pts)


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    sample = synthetic.make_fusion_sample(96, 128, 4, seed=0, batch=1)
    out = run(sample)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fusion_small.npz"),
                        cfg=np.array([96, 128, 4, 0, 1]), thresholds=np.array([*PROB_T, DISP_T, DEPTH_T, VIEW_T]),
                        **{k: v.numpy() for k, v in out.items()})
    print({k: (tuple(v.shape), float(v.float().mean())) for k, v in out.items()})
