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


def loss_handler():
    """The reference loss_handler.py over MagicMock stubs for its plotting / config imports.
    Several of its methods hard-code CUDA (`stroke_ids.cuda()`, `.to('cuda')`, `.get_device()`):
    use `cpu_cuda_shim()` around calls to run them on CPU tensors."""
    if "loss" not in _cache:
        from unittest.mock import MagicMock
        pytorch3d_chamfer()
        sys.modules["pytorch3d_chamfer"] = _cache["cham"]
        for name in ("omegaconf", "omegaconf.listconfig", "seaborn", "matplotlib", "matplotlib.pyplot", "matplotlib.path",
                     "matplotlib.patches", "point_cloud_utils", "pyvista", "wandb"):
            sys.modules.setdefault(name, MagicMock())
        if REF not in sys.path:
            sys.path.insert(0, REF)
        _cache["loss"] = _load("ref_loss_handler", os.path.join(REF, "loss_handler.py"))
    return _cache["loss"]


class cpu_cuda_shim:
    """Context manager: makes `.cuda()`, `.to('cuda', ...)` and `.get_device()` no-ops on CPU tensors so the
    reference's hard-coded device moves can be exercised without a GPU (test infrastructure only)."""

    def __enter__(self):
        import torch
        self._t = torch
        self._cuda, self._to, self._gd = torch.Tensor.cuda, torch.Tensor.to, torch.Tensor.get_device
        orig_to = self._to

        def to(self_, *a, **k):
            a = tuple(x for x in a if not (isinstance(x, str) and x.startswith("cuda")) and not (isinstance(x, int) and not isinstance(x, bool)))
            k = {kk: v for kk, v in k.items() if not (kk == "device" and str(v).startswith("cuda"))}
            return orig_to(self_, *a, **k) if (a or k) else self_

        torch.Tensor.cuda = lambda self_, *a, **k: self_
        torch.Tensor.to = to
        torch.Tensor.get_device = lambda self_: 0
        return self

    def __exit__(self, *exc):
        self._t.Tensor.cuda, self._t.Tensor.to, self._t.Tensor.get_device = self._cuda, self._to, self._gd
        return False
