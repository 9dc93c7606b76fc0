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


def load_devkit_data_classes():
    """Returns the reference's nuscenes-devkit/lyft_dataset_sdk/utils/data_classes.py module
    (LidarPointCloud.from_file / transform / remove_close).  pyquaternion is absent here and
    only needed by code paths that build matrices from quaternions, so it gets an empty shell;
    geometry_utils is executed from the reference too."""
    pq = _shell("pyquaternion")
    if not hasattr(pq, "Quaternion"):
        class Quaternion:  # only built as a default argument on the paths used
            def __init__(self, *a, **k):
                pass
        pq.Quaternion = Quaternion
    if "matplotlib" not in sys.modules:      # type annotations only (matplotlib.axes.Axes)
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            _shell("matplotlib")
            ax = _shell("matplotlib.axes")
            ax.Axes = object
    _shell("lyft_dataset_sdk")
    _shell("lyft_dataset_sdk.utils")
    base = os.path.join(REF, "nuscenes-devkit", "lyft_dataset_sdk", "utils")
    _exec("lyft_dataset_sdk.utils.geometry_utils", os.path.join(base, "geometry_utils.py"))
    return _exec("lyft_dataset_sdk.utils.data_classes", os.path.join(base, "data_classes.py"))


def run_second_get_sensor_points(info):
    """Executes the point-assembly statements of the reference's
    NuScenesDataset.get_sensor_data (second/second/data/nuscenes_dataset.py:196-223) on `info`
    (a dict with lidar_path, timestamp, sweeps[...] as the reference's info pickles hold) and
    returns its `points`.  The module itself cannot be imported (spconv, fire); the statements
    are read from the reference file at run time and executed unchanged."""
    import textwrap
    from pathlib import Path

    import numpy as np
    path = os.path.join(REF, "second", "second", "data", "nuscenes_dataset.py")
    with open(path) as f:
        lines = f.readlines()
    start = next(i for i, l in enumerate(lines) if "lidar_path = Path(info['lidar_path'])" in l)
    end = next(i for i, l in enumerate(lines) if "axis=0)[:, [0, 1, 2, 4]]" in l)
    src = textwrap.dedent("".join(lines[start:end + 1]))
    ns = {"np": np, "Path": Path, "info": info}
    exec(compile(src, path, "exec"), ns)
    return ns["points"]


def run_bev_draw_boxes(im, voxel_size, corners, class_ids, classes, z_offset=0.0):
    """Executes the reference's own ``draw_boxes`` (generating-dataset/generating_train_bev.py:127-139)
    together with the helpers it calls (:47-82) on `im` (H, W, 3) float32, in place.  They are nested
    closures of ``main``, so their source lines are read from the reference file,
    dedented and executed unchanged; only ``np.int0`` (removed in numpy 2) is aliased to ``np.intp``.
    The boxes are stand-ins that offer what draw_boxes reads: ``bottom_corners()`` and ``name``."""
    import textwrap
    import types

    import cv2
    import numpy as np
    path = os.path.join(REF, "generating-dataset", "generating_train_bev.py")
    with open(path) as f:
        lines = f.readlines()

    def block(name):
        start = next(i for i, l in enumerate(lines) if l.lstrip().startswith("def %s(" % name))
        indent = len(lines[start]) - len(lines[start].lstrip())
        end = start + 1
        while end < len(lines) and (not lines[end].strip() or len(lines[end]) - len(lines[end].lstrip()) > indent):
            end += 1
        return textwrap.dedent("".join(lines[start:end]))

    npx = types.SimpleNamespace(**{k: getattr(np, k) for k in dir(np) if not k.startswith("__")})
    npx.int0 = np.intp
    ns = {"np": npx, "cv2": cv2}
    for name in ("create_transformation_matrix_to_voxel_space", "transform_points", "car_to_voxel_coords",
                 "draw_boxes"):
        exec(compile(block(name), path, "exec"), ns)

    class _Box:
        def __init__(self, c, name):
            self._c, self.name = c, name

        def bottom_corners(self):
            return self._c

    boxes = [_Box(np.asarray(c, dtype=np.float64), classes[int(k)]) for c, k in zip(corners, class_ids)]
    ns["draw_boxes"](im, voxel_size, boxes=boxes, classes=classes, z_offset=z_offset)
    return im


def load_bev_closures():
    """The reference's own BEV closures (generating-dataset/generating_train_bev.py:47-104), nested in
    ``main``: their source lines are read from the reference file, dedented and executed
    unchanged; only ``np.int0`` (removed in numpy 2) is aliased to ``np.intp`` - the same C cast.
    Returns a dict name -> function."""
    import textwrap
    import types

    import numpy as np
    path = os.path.join(REF, "generating-dataset", "generating_train_bev.py")
    with open(path) as f:
        lines = f.readlines()

    def block(name):
        start = next(i for i, l in enumerate(lines) if l.lstrip().startswith("def %s(" % name))
        indent = len(lines[start]) - len(lines[start].lstrip())
        end = start + 1
        while end < len(lines) and (not lines[end].strip() or len(lines[end]) - len(lines[end].lstrip()) > indent):
            end += 1
        return textwrap.dedent("".join(lines[start:end]))

    npx = types.SimpleNamespace(**{k: getattr(np, k) for k in dir(np) if not k.startswith("__")})
    npx.int0 = np.intp
    ns = {"np": npx}
    names = ("create_transformation_matrix_to_voxel_space", "transform_points", "car_to_voxel_coords",
             "create_voxel_pointcloud", "normalize_voxel_intensities")
    for name in names:
        exec(compile(block(name), path, "exec"), ns)
    return {n: ns[n] for n in names}


def run_bev_closures(points, shape, voxel_size, z_offset):
    """points (3|4, N) float32 -> (raw counts f32, normalised f32, u8 image) exactly as
    prepare_training_data_for_scene computes them (generating_train_bev.py:210-213): the reference's
    create_voxel_pointcloud, normalize_voxel_intensities and the ``np.round(bev*255).astype(np.uint8)``
    statement."""
    import numpy as np
    fn = load_bev_closures()
    with np.errstate(invalid="ignore"):
        raw = fn["create_voxel_pointcloud"](points, shape, voxel_size=voxel_size, z_offset=z_offset)
    bev = fn["normalize_voxel_intensities"](raw)
    bev_im = np.round(bev * 255).astype(np.uint8)
    return raw, bev, bev_im


def _function_source(path, name):
    """Source of the top-level ``def name(`` of the reference file `path`, unchanged."""
    with open(path) as f:
        lines = f.readlines()
    start = next(i for i, l in enumerate(lines) if l.startswith("def %s(" % name))
    end = start + 1
    while end < len(lines) and (not lines[end].strip() or lines[end][0] in " \t"):
        end += 1
    return "".join(lines[start:end])


def load_collate_functions():
    """The reference's own batch assembly: ``merge_second_batch`` / ``merge_second_batch_multigpu``
    (second/second/data/preprocess.py:21-88) and ``example_convert_to_torch``
    (second/second/pytorch/train.py:34-62).  Neither module imports here (skimage, spconv, fire ...),
    so the three function bodies are read from the reference files and executed unchanged.
    Returns a dict name -> function."""
    from collections import defaultdict

    import numpy as np
    import torch
    pre = os.path.join(REF, "second", "second", "data", "preprocess.py")
    trn = os.path.join(REF, "second", "second", "pytorch", "train.py")
    ns = {"np": np, "defaultdict": defaultdict, "torch": torch}
    for path, name in ((pre, "merge_second_batch"), (pre, "merge_second_batch_multigpu"),
                       (trn, "example_convert_to_torch")):
        exec(compile(_function_source(path, name), path, "exec"), ns)
    return {n: ns[n] for n in ("merge_second_batch", "merge_second_batch_multigpu", "example_convert_to_torch")}
