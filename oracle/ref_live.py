"""The LIVE reference (TruongKhang/cds-mvsnet, ``models/`` package, unmodified) as shipped to ``oracle/_ref/``.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Nothing in the product path (``cds_mvsnet_b200/``) imports this module; the callers
are ``tests/`` (live parity and live ``patch()`` tests), ``bench.py --impl reference`` (the reference's own CPU forward) and
``bench.py``'s ``reference_eager_gpu`` leg (the incumbent: the same unmodified code run eagerly on the same B200).

``oracle/_ref/`` is git-ignored and filled by ``__graft_entry__.build()`` in the build container (recipe: ``ship_reference``
below): the reference's ``models`` package, verbatim, as ONE zip archive that Python imports in place (zipimport) -- it
travels to the GPU box with the gpurun snapshot like the built ``.so`` files.  No reference source is committed to this
repository and none is unpacked into its tree.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_ZIP = os.path.join(REF_DIR, "reference_models.zip")
SRC = os.environ.get("CDS_REF_PATH", "/root/reference")


def ship_reference(verbose: bool = True) -> bool:
    """Archive the reference's ``models`` package (7 .py files, no weights) as oracle/_ref/reference_models.zip.  Build
    container only: on the GPU box /root/reference does not exist and the archive that travelled is used as is."""
    src = os.path.join(SRC, "models")
    if not os.path.isdir(src):
        return available()
    os.makedirs(REF_DIR, exist_ok=True)
    tmp = REF_ZIP + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for root, dirs, files in os.walk(src):
            dirs[:] = sorted(d for d in dirs if d != "__pycache__")
            for f in sorted(files):
                if f.endswith(".py"):
                    full = os.path.join(root, f)
                    z.write(full, os.path.join("models", os.path.relpath(full, src)))
        meta = os.path.join(SRC, ".SUBMODULES.json")
        z.writestr("SOURCE.txt", f"verbatim archive of {src} made by oracle/ref_live.py:ship_reference (git-ignored)\n" +
                   (open(meta).read() if os.path.exists(meta) else ""))
    os.replace(tmp, REF_ZIP)
    if verbose:
        print(f"[oracle/_ref] reference models/ archived from {src}")
    return True


def available() -> bool:
    return os.path.isfile(REF_ZIP)


def load():
    """(models.model, models.module, models.dynamic_conv, models.utils.warping) of the shipped reference."""
    if not available():
        raise RuntimeError("oracle/_ref/reference_models.zip is missing: run __graft_entry__.build() in the build container "
                           "(it archives /root/reference/models there; the archive travels to the GPU box)")
    if REF_ZIP not in sys.path:
        sys.path.insert(0, REF_ZIP)
    with contextlib.redirect_stdout(io.StringIO()):
        import models.dynamic_conv as rdyn
        import models.model as rmodel
        import models.module as rmodule
        import models.utils.warping as rwarp
    if not os.path.abspath(rmodel.__file__).startswith(REF_ZIP):
        raise RuntimeError(f"a different 'models' package is already imported from {rmodel.__file__}")
    return rmodel, rmodule, rdyn, rwarp


def build_model(state_dict, ndepths, ratios, refine=False, device="cpu", rmodel=None):
    """The reference's own ``models.model.CDSMVSNet`` (as bound in ``rmodel`` -- patched or not) with ``state_dict`` loaded."""
    import torch
    rmodel = rmodel or load()[0]
    with contextlib.redirect_stdout(io.StringIO()):   # the constructor prints banners
        m = rmodel.CDSMVSNet(refine=refine, ndepths=tuple(ndepths), depth_interals_ratio=tuple(ratios), share_cr=False,
                             cr_base_chs=(8,) * len(ndepths), grad_method="detach")
    own = m.state_dict()
    m.load_state_dict({k: v for k, v in state_dict.items() if k in own}, strict=True)
    return m.to(torch.device(device)).eval()
