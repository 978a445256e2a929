"""Imports the REAL reference from /root/reference -- TEST INFRASTRUCTURE, this container only.

/root/reference does not exist on the GPU box; nothing under tests/ -m gpu, smoke() or bench.py may
call this module.  It is used by oracle/make_golden.py (fixture generation) and by the CPU tests
that are skipped when the reference tree is absent.

Recipe: the encoder files need only torch+numpy and are loaded by path under
a stub package (models/__init__.py drags in omegaconf); pytorch3d_chamfer.py is imported
UNMODIFIED over a sys.modules shim that supplies pytorch3d.ops.knn.{knn_points,knn_gather} from
oracle.torch_oracle (pytorch3d itself is not installable offline).
"""
import importlib.util
import os
import sys
import types

REF = os.environ.get("MASKPLANNER_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF, "models", "pointnet2_utils.py"))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def pointnet2_utils():
    """The reference models/pointnet2_utils.py module object."""
    if "pn2" not in _cache:
        sys.dont_write_bytecode = True
        pkg = types.ModuleType("refmodels")
        pkg.__path__ = [os.path.join(REF, "models")]
        sys.modules["refmodels"] = pkg
        _cache["pn2"] = _load("refmodels.pointnet2_utils", os.path.join(REF, "models", "pointnet2_utils.py"))
    return _cache["pn2"]


def pointnet2_cls_ssg():
    """The reference models/pointnet2_cls_ssg.py (its relative import resolves to the module above)."""
    if "ssg" not in _cache:
        pointnet2_utils()
        _cache["ssg"] = _load("refmodels.pointnet2_cls_ssg", os.path.join(REF, "models", "pointnet2_cls_ssg.py"))
    return _cache["ssg"]


def pytorch3d_chamfer():
    """The reference pytorch3d_chamfer.py, unmodified, over the knn shim."""
    if "cham" not in _cache:
        sys.dont_write_bytecode = True
        from . import torch_oracle as T

        class Pointclouds:  # the reference only uses it in isinstance() checks
            pass

        for name in ("pytorch3d", "pytorch3d.ops", "pytorch3d.ops.knn", "pytorch3d.structures",
                     "pytorch3d.structures.pointclouds"):
            sys.modules.setdefault(name, types.ModuleType(name))
        sys.modules["pytorch3d.ops.knn"].knn_points = T.knn_points
        sys.modules["pytorch3d.ops.knn"].knn_gather = T.knn_gather
        sys.modules["pytorch3d.structures.pointclouds"].Pointclouds = Pointclouds
        _cache["cham"] = _load("ref_pytorch3d_chamfer", os.path.join(REF, "pytorch3d_chamfer.py"))
    return _cache["cham"]
