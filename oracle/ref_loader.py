"""Loads the reference's own hot-path files from /root/reference for golden-vector
generation.  TEST INFRASTRUCTURE; only usable in the build container
(/root/reference does not exist on the GPU box) and only called by
oracle/gen_golden.py and by tests that skip when the reference is absent.

Normal package import of ``second`` fails here (torchplus imports
``collections.Iterable``; box_np_ops imports spconv - SURVEY.md 8c), so the three
files are executed individually with empty package shells:

  second/second/pytorch/models/voxel_encoder.py   (registry + get_paddings_indicator)
  second/second/pytorch/models/pointpillars.py    (PillarFeatureNet*, PointPillarsScatter)
  second/second/utils/simplevis.py                (points_to_bev, sibling of points_to_voxel)
"""
import importlib.util
import os
import sys
import types

REF = os.environ.get("LYFT_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "second", "second"))


def _shell(name):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    return sys.modules[name]


def _exec(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_pointpillars():
    """Returns (pointpillars module, voxel_encoder module) of the reference."""
    import torch

    for n in ("second", "second.pytorch", "second.pytorch.models", "torchplus"):
        _shell(n)
    tp_nn = _shell("torchplus.nn")

    class Empty(torch.nn.Module):  # stand-in for torchplus.nn.Empty (identity)
        def __init__(self, *a, **k):
            super().__init__()

        def forward(self, *args, **kwargs):
            return args[0] if len(args) == 1 else args

    tp_nn.Empty = Empty
    tp_nn.GroupNorm = torch.nn.GroupNorm
    tp_nn.Sequential = torch.nn.Sequential
    tools = _exec("torchplus.tools", os.path.join(REF, "second", "torchplus", "tools.py"))
    assert hasattr(tools, "change_default_args")
    mid = _shell("second.pytorch.models.middle")
    registry = {}

    def register_middle(cls, name=None):
        registry[name or cls.__name__] = cls
        return cls

    mid.register_middle = register_middle
    mid.REGISTERED_MIDDLE_CLASSES = registry
    base = os.path.join(REF, "second", "second", "pytorch", "models")
    ve = _exec("second.pytorch.models.voxel_encoder", os.path.join(base, "voxel_encoder.py"))
    pp = _exec("second.pytorch.models.pointpillars", os.path.join(base, "pointpillars.py"))
    return pp, ve


def load_simplevis():
    """Returns the reference's second/second/utils/simplevis.py module."""
    _shell("second")
    core = _shell("second.core")
    if not hasattr(core, "box_np_ops"):
        core.box_np_ops = _shell("second.core.box_np_ops")  # only used by drawing helpers
    return _exec("second.utils.simplevis_ref",
                 os.path.join(REF, "second", "second", "utils", "simplevis.py"))
