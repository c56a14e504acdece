"""cds_mvsnet_b200 -- the CDS-MVSNet depth-inference hot path as hand-written sm_100a CUDA kernels
behind the reference's own Python call surface.

    from cds_mvsnet_b200 import CDSMVSNet           # drop-in for models.model.CDSMVSNet (eval; refine=False and refine=True)
    from cds_mvsnet_b200 import patch, unpatch      # rebind the hot-path names inside the reference's modules (4 levels)

Importing the package never touches the GPU; the first op call loads ``libcds_b200.so`` (built by
``python -m cds_mvsnet_b200.build``) and raises if it or a B200 is missing -- there is no fallback.
"""
from .modules import (CDSMVSNet, CostRegNet, DynamicConv, FeatureNet, Refinement, StageNet, conf_regression,  # noqa: F401
                      depth_regression, homo_warping_3D, patch, unpatch, PATCH_LEVELS)

__version__ = "0.1.0"
