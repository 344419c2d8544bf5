"""TEST INFRASTRUCTURE ONLY — loader that imports the *unmodified* reference modules
from /root/reference/code on a CPU-only box.

Only usable in the build container (the GPU box has no /root/reference).  Used by
oracle/gen_golden.py to produce tests/golden/*.npz and by tests that pin oracle/port.py
against the reference when the reference tree is present.

What is shimmed:
  * ``torch.Tensor.cuda`` / ``nn.Module.cuda`` -> identity (every hot-path module
    hard-codes ``.cuda()``: density.py:18, ray_sampler.py:23.., rend_util.py:58..).
  * stub modules for packages that are absent here and not needed by the eval path:
    hydra, kaolin, nerfacc, trimesh, imageio, skimage, pytorch3d.  ``pytorch3d.ops.knn_points``
    and the three nerfacc functions are supplied from oracle/port.py (restatements,
    "parity unpinned" at that boundary — the packages are unpinned and absent).
  * ``lib.model.smpl.SMPLServer`` needs the licence-gated SMPL pkl -> replaced by a stub
    class; deformers are built with explicit verts/weights.
"""
import sys
import types
import importlib

REF_CODE = "/root/reference/code"


def available():
    import os
    return os.path.isdir(REF_CODE + "/lib/model")


class AttrDict(dict):
    """OmegaConf stand-in: attribute access + .get()."""
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return AttrDict(v) if isinstance(v, dict) and not isinstance(v, AttrDict) else v

    def __setattr__(self, k, v):
        self[k] = v


_loaded = {}


def load():
    """Returns a namespace with the reference modules (networks, density, ray_sampler,
    deformer, embedders, rend_util, multiply)."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError("reference tree not present (only exists in the build container)")
    import torch
    from . import port

    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
        return m

    hydra = stub("hydra")
    hydra.utils = stub("hydra.utils", to_absolute_path=lambda p: p)
    hydra.main = lambda *a, **k: (lambda f: f)
    kaolin = stub("kaolin")
    kaolin.ops = stub("kaolin.ops")
    kaolin.ops.mesh = stub("kaolin.ops.mesh", index_vertices_by_faces=lambda v, f: None,
                           check_sign=None)
    kaolin.metrics = stub("kaolin.metrics")
    kaolin.metrics.trianglemesh = stub("kaolin.metrics.trianglemesh")
    stub("nerfacc", render_weight_from_density=port.render_weight_from_density,
         pack_info=port.pack_info, accumulate_along_rays=port.accumulate_along_rays)
    tm = stub("trimesh")
    tm.ray = stub("trimesh.ray")
    tm.ray.ray_triangle = stub("trimesh.ray.ray_triangle")
    tm.primitives = stub("trimesh.primitives")
    stub("imageio")
    sk = stub("skimage")
    sk.measure = stub("skimage.measure")
    p3d = stub("pytorch3d")
    p3d.ops = stub("pytorch3d.ops", knn_points=port.knn_points)
    try:
        import cv2  # noqa: F401
    except Exception:
        stub("cv2")

    if REF_CODE not in sys.path:
        sys.path.insert(0, REF_CODE)

    # SMPLServer needs the SMPL pkl: stub the module before deformer/multiply import it.
    smpl_stub = types.ModuleType("lib.model.smpl")

    class SMPLServer(torch.nn.Module):  # never instantiated by the oracle
        def __init__(self, *a, **k):
            super().__init__()
            raise RuntimeError("SMPL model files are licence-gated and absent")
    smpl_stub.SMPLServer = SMPLServer
    importlib.import_module("lib")
    importlib.import_module("lib.model")
    sys.modules["lib.model.smpl"] = smpl_stub

    ns = types.SimpleNamespace()
    ns.networks = importlib.import_module("lib.model.networks")
    ns.density = importlib.import_module("lib.model.density")
    ns.embedders = importlib.import_module("lib.model.embedders")
    ns.ray_sampler = importlib.import_module("lib.model.ray_sampler")
    ns.deformer = importlib.import_module("lib.model.deformer")
    ns.rend_util = importlib.import_module("lib.utils.rend_util")
    ns.multiply = importlib.import_module("lib.model.multiply")
    ns.lbs = importlib.import_module("lib.smpl.lbs")
    ns.sampler_cls = importlib.import_module("lib.model.sampler").PointInSpace       # multiply.py:67
    ns.AttrDict = AttrDict
    _loaded["ns"] = ns
    return ns
