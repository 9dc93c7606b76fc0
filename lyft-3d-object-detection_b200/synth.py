"""Seeded synthetic Lyft-shaped inputs.

Host-side (numpy) definitions used by the parity tests and by bench.py.  Every
cloud derives from the one real sweep that ships with the reference
(host-a011_lidar1_1233090652702363606.bin, 53,146 points, 5 float32 per point:
x, y, z, intensity, ring - lyft_dataset_sdk/utils/data_classes.py:269-284), a
copy of which is committed under tests/golden/.
"""
import os

import numpy as np

FIXTURE_NAME = "host-a011_lidar1_1233090652702363606.bin"
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE_PATH = os.path.join(_ROOT, "tests", "golden", FIXTURE_NAME)

# C1 - generating-dataset/generating_train_bev.py:37-39
BEV_SHAPE = (336, 336, 3)
BEV_VOXEL_SIZE = (0.4, 0.4, 1.5)
BEV_Z_OFFSET = -2.0
# C2 - second/second/configs/all.fhd.config:4-9,296
SECOND_VOXEL_SIZE = (0.05, 0.05, 0.1)
SECOND_RANGE = (0.0, -32.0, -3.0, 52.8, 32.0, 1.0)
SECOND_MAX_POINTS = 5
SECOND_MAX_VOXELS = 60000
# C3 - second/second/configs/nuscenes/all.pp.largea.config:4-9,278,358
PILLAR_VOXEL_SIZE = (0.25, 0.25, 20.0)
PILLAR_RANGE = (-50.0, -50.0, -10.0, 50.0, 50.0, 10.0)
PILLAR_MAX_POINTS = 60
PILLAR_MAX_VOXELS = 30000
PILLAR_CANVAS = (400, 400)  # ny, nx
PILLAR_FEATURES = 64
# C4
BEV1024_SHAPE = (1024, 1024, 3)
BEV1024_VOXEL_SIZE = (0.2, 0.2, 1.5)


def load_fixture_raw(path=None):
    """(N,5) float32 rows as stored in the .bin."""
    scan = np.fromfile(path or FIXTURE_PATH, dtype=np.float32)
    return scan.reshape(-1, 5)


def fixture_points_4xn(path=None):
    """(4,N) float32 - the LidarPointCloud.points layout (data_classes.py:282-284)."""
    return load_fixture_raw(path)[:, :4].T


def fixture_points_nx4(path=None):
    """(N,4) float32 C-contiguous x,y,z,intensity."""
    return np.ascontiguousarray(load_fixture_raw(path)[:, :4])


def multisweep_cloud(n_copies=11, base=None):
    """C2/C3 cloud: ``n_copies`` jittered copies of the fixture in sweep order.

    copy s: xyz += N(0, 0.02 m) from default_rng(1000+s) (float32 add),
    x += 0.5*s, channel 3 = 0.05*s (time lag).  11 copies -> 584,606 points,
    20 copies -> 1,062,920 points.  Returns (N,4) float32 C-contiguous.
    """
    base = fixture_points_nx4() if base is None else base
    out = np.empty((n_copies * base.shape[0], 4), dtype=np.float32)
    n = base.shape[0]
    for s in range(n_copies):
        rng = np.random.default_rng(1000 + s)
        noise = rng.normal(0.0, 0.02, size=(n, 3)).astype(np.float32)
        blk = out[s * n:(s + 1) * n]
        blk[:, :3] = base[:, :3] + noise
        blk[:, 0] += np.float32(0.5 * s)
        blk[:, 3] = np.float32(0.05 * s)
    return out


def c5_frame(f, base=None):
    """C5 frame f: fixture rotated about z by 2*pi*u_f plus N(0,0.02) jitter,
    u_f and noise from default_rng(5000+f).  (N,4) float32, channel 3 = intensity."""
    base = fixture_points_nx4() if base is None else base
    rng = np.random.default_rng(5000 + f)
    u = rng.random()
    a = 2.0 * np.pi * u
    c, s = np.float32(np.cos(a)), np.float32(np.sin(a))
    noise = rng.normal(0.0, 0.02, size=(base.shape[0], 3)).astype(np.float32)
    out = np.empty_like(base)
    out[:, 0] = c * base[:, 0] - s * base[:, 1] + noise[:, 0]
    out[:, 1] = s * base[:, 0] + c * base[:, 1] + noise[:, 1]
    out[:, 2] = base[:, 2] + noise[:, 2]
    out[:, 3] = base[:, 3]
    return out


def c5_frames(frame_ids, base=None):
    """Concatenated C5 frames + int64 frame offsets (len F+1)."""
    base = fixture_points_nx4() if base is None else base
    frames = [c5_frame(int(f), base) for f in frame_ids]
    offs = np.zeros(len(frames) + 1, dtype=np.int64)
    offs[1:] = np.cumsum([fr.shape[0] for fr in frames])
    return np.concatenate(frames, axis=0), offs


def sweep_transform(s):
    """C4 per-sweep 4x4 (float64): yaw 0.01*s rad about z, t = (0.5*s, 0, 0)."""
    a = 0.01 * s
    tm = np.eye(4, dtype=np.float64)
    tm[0, 0], tm[0, 1], tm[1, 0], tm[1, 1] = np.cos(a), -np.sin(a), np.sin(a), np.cos(a)
    tm[0, 3] = 0.5 * s
    return tm


def map_raster(shape_hw=(1024, 1024), seed=4000, blobs=40):
    """C4 map raster: seeded uint8 (H,W,3) with 0/255 rectangular blobs."""
    rng = np.random.default_rng(seed)
    h, w = shape_hw
    m = np.zeros((h, w, 3), dtype=np.uint8)
    for _ in range(blobs):
        y0, x0 = int(rng.integers(0, h)), int(rng.integers(0, w))
        dy, dx = int(rng.integers(8, h // 4)), int(rng.integers(8, w // 4))
        ch = int(rng.integers(0, 3))
        m[y0:y0 + dy, x0:x0 + dx, ch] = 255
    return m


def ingest_case(n_sweeps=4, n_key=6000, seed=77):
    """Seeded multi-sweep case for the ingest path: slices of the bundled
    sweep as raw (M,5) rows, a full 3-D rotation + translation per sweep (float64) and
    microsecond timestamps - the fields of the reference's info["sweeps"] entries."""
    rng = np.random.default_rng(seed)
    raw = load_fixture_raw()
    raw = raw.copy()
    raw[::50, :2] *= 0.01          # a few returns close to the sensor (exercise remove_close)
    key = np.ascontiguousarray(raw[:n_key])
    ts_us = 1_550_000_000_000_000
    sweeps = []
    at = n_key
    for s in range(n_sweeps):
        m = 3000 + 517 * s
        ang = rng.normal(size=3) * 0.05
        cx, cy, cz = np.cos(ang)
        sx, sy, sz = np.sin(ang)
        rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
        sweeps.append({"points": np.ascontiguousarray(raw[at:at + m]),
                       "sweep2lidar_rotation": rz @ ry @ rx,
                       "sweep2lidar_translation": rng.normal(size=3) * np.array([2.0, 0.5, 0.05]),
                       "timestamp": ts_us - 50_000 * (s + 1)})
        at += m
    return key, sweeps, ts_us


# Lyft classes (generating-dataset/generating_train_bev.py:34-35) and typical footprints (w, l) in metres
BOX_CLASSES = ["car", "motorcycle", "bus", "bicycle", "truck", "pedestrian", "other_vehicle", "animal",
               "emergency_vehicle"]
_BOX_WL = [(1.93, 4.76), (0.96, 2.35), (2.96, 12.34), (0.63, 1.76), (2.84, 10.24), (0.77, 0.81), (2.79, 8.20),
           (0.36, 0.73), (2.45, 6.52)]
BOX_SCALE = 0.8   # generating_train_bev.py:40


def box_scene(seed, n_boxes=60, extent=75.0):
    """Seeded annotation set of one sample in car space: (corners, class_ids) with corners (n, 3, 4)
    float64 = Box.bottom_corners() of every box (four bottom corners in order around the footprint,
    box sizes already scaled by BOX_SCALE, generating_train_bev.py:119-125) and class_ids (n,) the index
    into BOX_CLASSES.  Centres reach beyond the 336 x 0.4 m image so that some boxes are clipped."""
    rng = np.random.default_rng(seed)
    cls = rng.integers(0, len(BOX_CLASSES), n_boxes)
    corners = np.zeros((n_boxes, 3, 4), dtype=np.float64)
    for i, c in enumerate(cls):
        w, l = np.array(_BOX_WL[c]) * BOX_SCALE * rng.uniform(0.8, 1.25)
        yaw = rng.uniform(-np.pi, np.pi)
        centre = rng.uniform(-extent, extent, 2)
        z = rng.uniform(-2.5, -0.5)
        local = np.array([[l / 2, w / 2], [l / 2, -w / 2], [-l / 2, -w / 2], [-l / 2, w / 2]])
        rot = np.array([[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]])
        xy = local @ rot.T + centre
        corners[i, 0], corners[i, 1], corners[i, 2] = xy[:, 0], xy[:, 1], z
    return corners, cls.astype(np.int32)
