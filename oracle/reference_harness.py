"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference from /root/reference.

This file exists to pin `oracle/modest_oracle.py` (the travelling CPU restatement) against
the reference's own code, and to generate the golden vectors under tests/golden/.  It can
only run in the build container (the GPU box has no /root/reference); nothing in the
product, in `-m gpu` tests, in smoke() or in bench.py imports it.

The reference needs four modules that are not installed here:
`hydra`, `omegaconf`, `pyquaternion`, `iou3d_nms_cuda`.  They are replaced by minimal stand-ins
*before* the reference modules are imported; the reference's own numerics are untouched.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("MODEST_REFERENCE_ROOT", "/root/reference")
_GCM = os.path.join(REFERENCE_ROOT, "generate_cluster_mask")


def available() -> bool:
    return os.path.isdir(_GCM)


class _AttrDict(dict):
    """Enough of omegaconf.DictConfig for the reference's `args.a.b` / `args.get()` / `**args.x`."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return _AttrDict(v) if isinstance(v, dict) and not isinstance(v, _AttrDict) else v

    def __setattr__(self, k, v):
        self[k] = v


def _install_stubs():
    if "hydra" not in sys.modules:
        hydra = types.ModuleType("hydra")
        hydra.main = lambda **kw: (lambda f: f)
        sys.modules["hydra"] = hydra
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")
        oc.DictConfig = _AttrDict

        class OmegaConf:  # noqa: D401 - stand-in
            @staticmethod
            def to_yaml(cfg):
                import yaml
                return yaml.safe_dump(dict(cfg))

            @staticmethod
            def save(config, f):
                import yaml
                with open(f, "w") as fh:
                    yaml.safe_dump(dict(config), fh)

        oc.OmegaConf = OmegaConf
        sys.modules["omegaconf"] = oc
    if "pyquaternion" not in sys.modules:
        pq = types.ModuleType("pyquaternion")

        class Quaternion:
            """pyquaternion's axis-angle constructor and transformation_matrix, restated
            (q-matrix times conjugate q-bar-matrix, rows/cols 1..3)."""

            def __init__(self, axis, angle):
                ax = np.asarray(axis, float)
                ax = ax / np.linalg.norm(ax)
                self.q = np.concatenate([[np.cos(angle / 2.0)], ax * np.sin(angle / 2.0)])

            @property
            def transformation_matrix(self):
                w, x, y, z = self.q
                qm = np.array([[w, -x, -y, -z], [x, w, -z, y], [y, z, w, -x], [z, -y, x, w]])
                qb = np.array([[w, -x, -y, -z], [x, w, z, -y], [y, -z, w, x], [z, y, -x, w]])
                m = np.eye(4)
                m[:3, :3] = np.dot(qm, qb.conj().transpose())[1:][:, 1:]
                return m

        pq.Quaternion = Quaternion
        sys.modules["pyquaternion"] = pq
    if "iou3d_nms_cuda" not in sys.modules:
        ref_so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
        if os.path.isdir(ref_so) and ref_so not in sys.path:
            sys.path.insert(0, ref_so)
        try:
            import torch  # noqa: F401  (the extension links against libtorch)
            import iou3d_nms_cuda  # noqa: F401  built by oracle/build_ref.py
        except Exception:
            sys.modules["iou3d_nms_cuda"] = types.ModuleType("iou3d_nms_cuda")


_loaded = None


def load():
    """Return a namespace with the reference's hot-path modules."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    # the reference imports `utils.*` relative to generate_cluster_mask/
    for name in [m for m in sys.modules if m == "utils" or m.startswith("utils.")]:
        del sys.modules[name]
    sys.path.insert(0, _GCM)
    try:
        import pre_compute_pp_score as ref_pp          # noqa
        from utils import pointcloud_utils as ref_pc    # noqa
        from utils import clustering_utils as ref_cl    # noqa
        from utils import kitti_util as ref_ku          # noqa
        import generate_mask as ref_gm                  # noqa
        import combine_labels as ref_cb                 # noqa
    finally:
        sys.path.remove(_GCM)
    ns = types.SimpleNamespace(pp=ref_pp, pc=ref_pc, cl=ref_cl, ku=ref_ku, gm=ref_gm, cb=ref_cb, AttrDict=_AttrDict)
    # keep the reference's `utils` package from shadowing anything of ours
    ns._mods = {k: sys.modules.pop(k) for k in list(sys.modules)
                if k == "utils" or k.startswith("utils.")}
    _loaded = ns
    return ns
